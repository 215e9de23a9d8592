"""Run an UNMODIFIED reference driver script with its LrgNet step executed by the sm_100a engine:

    python -m learn_region_grow_b200.run_reference /path/to/learn_region_grow/test_region_grow.py --area 5

The drop-in directory goes in front of ``sys.path`` so ``import tensorflow`` / ``from learn_region_grow_util import *``
inside the script resolve to learn_region_grow_b200/dropin; the script's own directory follows (for ``class_util``),
and the h5py / matplotlib stand-ins are appended last so real installations win.

``--py2-range`` (first argument) gives the script Python 2's list-returning ``range`` as a module global: the reference's
test_beam_search.py concatenates ``range(n) + list(...)`` (:212,224) and runs unchanged that way.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def py2_range(*a):
    return list(range(*a))


def run(script, argv, init_globals=None):
    dropin = os.path.join(HERE, 'dropin')
    old_path, old_argv = list(sys.path), list(sys.argv)
    sys.path[:0] = [dropin, os.path.dirname(HERE), os.path.dirname(os.path.abspath(script))]
    sys.path.append(os.path.join(dropin, 'standins'))
    sys.argv = [script] + list(argv)
    try:
        return runpy.run_path(script, init_globals=init_globals, run_name='__main__')
    finally:
        sys.path[:] = old_path
        sys.argv = old_argv


if __name__ == '__main__':
    args = sys.argv[1:]
    g = None
    if args and args[0] == '--py2-range':
        g, args = {'range': py2_range}, args[1:]
    run(args[0], args[1:], g)
