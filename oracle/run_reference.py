"""Run an UNMODIFIED reference driver script on the host with the CPU oracle forward -- TEST INFRASTRUCTURE.

    python -m oracle.run_reference /root/reference/test_region_grow.py --area 5

sys.path order: oracle/ref_shims (CPU ``learn_region_grow_util``), learn_region_grow_b200/dropin (session-API
stand-in ``tensorflow``), the reference directory (for its own ``class_util``), and last the h5py / matplotlib
stand-ins (only reached when the real packages are absent).  ``runpy.run_path`` does not put the
script's directory first, so the reference's TF-based util is never imported.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def shim_paths(ref_dir):
    dropin = os.path.join(REPO, 'learn_region_grow_b200', 'dropin')
    # ref_shims first: its CPU learn_region_grow_util shadows the CUDA one in dropin/
    head = [os.path.join(HERE, 'ref_shims'), REPO, dropin, ref_dir]
    tail = [os.path.join(dropin, 'standins')]      # h5py / matplotlib stand-ins only if the real ones are absent
    return head, tail


def py2_range(*a):
    """Python 2's ``range``: a list.  test_beam_search.py:212,224 concatenate ``range(n) + list(...)``, which only Python 2
    accepts; handing this to the script as a module global runs it unmodified."""
    return list(range(*a))


def run(script, argv, init_globals=None):
    head, tail = shim_paths(os.path.dirname(os.path.abspath(script)))
    old_path, old_argv = list(sys.path), list(sys.argv)
    sys.path[:0] = head
    sys.path.extend(tail)
    sys.argv = [script] + list(argv)
    try:
        return runpy.run_path(script, init_globals=init_globals, run_name='__main__')
    finally:
        sys.path[:] = old_path
        sys.argv = old_argv


if __name__ == '__main__':
    run(sys.argv[1], sys.argv[2:])
