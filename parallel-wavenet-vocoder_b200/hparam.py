"""Config singleton of the generation path.

Mirrors the reference's config surface (reference hparam.py:7-69): two multi-document YAML
files, `hparams/default.yaml` and `hparams/hparams.yaml`; the documents of each file are
flattened into one mapping; the mapping stored under the case name in the user file is merged
over the defaults (user value wins, recursion only where both sides are mappings, reference
hparam.py:17-24); the result is exposed through a process-wide attribute-access object `hparam`
and `hparam.logdir` is derived as `<logdir_path>/<case>` (reference hparam.py:64-68).

Differences, all deliberate: `yaml.safe_load_all` (the reference's loader-less `yaml.load_all`
raises on PyYAML >= 6), files are resolved against the CWD first (as the reference does) and
then against this repository, and unknown attributes raise AttributeError instead of KeyError.
"""
import os

import yaml

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _resolve(path):
    if os.path.isabs(path) or os.path.exists(path):
        return path
    alt = os.path.join(_REPO_ROOT, path)
    return alt if os.path.exists(alt) else path


def read_yaml_documents(path):
    """All documents of a YAML file flattened into one dict (later documents win)."""
    flat = {}
    with open(_resolve(path), 'r') as fh:
        for document in yaml.safe_load_all(fh):
            if document:
                flat.update(document)
    return flat


def overlay(user, default):
    """Recursive user-over-default merge; returns `user` updated in place when it is a mapping."""
    if not (isinstance(user, dict) and isinstance(default, dict)):
        return user
    for key, dvalue in default.items():
        user[key] = overlay(user[key], dvalue) if key in user else dvalue
    return user


class AttrDict(dict):
    """dict with attribute access; nested mappings are converted on construction."""

    def __init__(self, mapping=None):
        super().__init__()
        for key, value in (mapping or {}).items():
            self[key] = AttrDict(value) if isinstance(value, dict) else value

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


class Hparam(AttrDict):
    def set_hparam_yaml(self, case, default_file='hparams/default.yaml', user_file='hparams/hparams.yaml'):
        defaults = read_yaml_documents(default_file)
        cases = read_yaml_documents(user_file)
        merged = overlay(cases[case], defaults) if case in cases else defaults
        self.clear()
        self.update(AttrDict(merged))
        self['case'] = case
        self['logdir'] = '{}/{}'.format(self['logdir_path'], case)
        return self

    def set_hparam_dict(self, mapping, case='inline'):
        """Populate from an in-memory mapping merged over default.yaml (tests / bench)."""
        merged = overlay(dict(mapping), read_yaml_documents('hparams/default.yaml'))
        self.clear()
        self.update(AttrDict(merged))
        self['case'] = case
        self['logdir'] = '{}/{}'.format(self['logdir_path'], case)
        return self


hparam = Hparam()
