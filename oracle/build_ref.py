"""Compile the reference's own CUDA kernels, UNMODIFIED and from where they lie under /root/reference, into
oracle/_ref/ -- TEST INFRASTRUCTURE (oracle/__init__.py).  Nothing is copied into the repo; oracle/_ref is git-ignored
but travels to the GPU box with the snapshot, where tests/test_tfops_gpu.py calls the C++-mangled ``*Launcher`` symbols
(tf_sampling_g.cu:194-211, tf_grouping_g.cu:125-141) to pin our kernels and the C restatement against them.

tf_ops/3d_interpolation/tf_interpolate.cpp includes four TensorFlow headers (:6-9) and TensorFlow is absent; it is
compiled UNMODIFIED all the same, with g++, against the from-scratch stand-in headers of oracle/tf_stubs (op registration
and a tiny OpKernel / Tensor surface) plus oracle/tf_stubs/harness.cpp (C ABI) -> oracle/_ref/libref_interpolate.so: the
reference's own three_nn / three_interpolate loops and its OpKernels' shape checks then pin oracle/tfops_oracle.c
(tests/test_cpu_suite.py) and the CUDA kernels (tests/test_tfops_gpu.py).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
OUT = os.path.join(HERE, '_ref')
UNITS = {
    'libref_sampling.so': 'tf_ops/sampling/tf_sampling_g.cu',
    'libref_grouping.so': 'tf_ops/grouping/tf_grouping_g.cu',
}


def build_oracle():
    subprocess.run(['make', '-s', '-C', HERE], check=True)
    return os.path.join(HERE, '_build', 'liboracle_tfops.so')


def build_ref():
    if not os.path.isdir(REF):
        return None          # GPU box: use the prebuilt files
    os.makedirs(OUT, exist_ok=True)
    for lib, src in UNITS.items():
        dst = os.path.join(OUT, lib)
        srcp = os.path.join(REF, src)
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(srcp):
            continue
        subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC', '-x', 'cu',
                        srcp, '-o', dst], check=True)
    dst = os.path.join(OUT, 'libref_interpolate.so')
    srcs = [os.path.join(HERE, 'tf_stubs', 'harness.cpp'), os.path.join(REF, 'tf_ops/3d_interpolation/tf_interpolate.cpp')]
    deps = srcs + [os.path.join(HERE, 'tf_stubs', 'tensorflow', 'core', 'framework', 'op_kernel.h')]
    if not os.path.exists(dst) or any(os.path.getmtime(dst) < os.path.getmtime(d) for d in deps):
        subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-std=c++14', '-I', os.path.join(HERE, 'tf_stubs')] + srcs + ['-o', dst], check=True)
    return OUT


if __name__ == '__main__':
    print(build_oracle())
    print(build_ref())
