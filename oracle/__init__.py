"""CPU oracle of the reference generation path -- test infrastructure only (see iaf_oracle.py)."""
