#!/usr/bin/env python
"""`python generate_multi.py case1 case2 ...` (reference generate_multi.py:13-24)."""
import sys

from generate import generate


def generate_multi(*cases):
    for case in cases:
        generate(case)
        print("case '{}' Done.".format(case))
    print('Done.')


if __name__ == '__main__':
    generate_multi(*sys.argv[1:])
