"""Build recipes for the oracle's native pieces (TEST INFRASTRUCTURE).

build_c():   gcc -> oracle/libdcn_ref.so   (plain-C DCNv2 restatement, oracle/dcn_ref.c)
build_ref(): compiles the REFERENCE's own DCN CUDA extension, from its sources where they lie
             under /root/reference (never copied into this repo), into oracle/_ref/ so that the
             GPU box can run the real reference kernels as a second DCN oracle and as the
             in-tree GPU competitor.  Only possible where /root/reference exists (the build
             container); the GPU box just uses the prebuilt .so that travels with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/basicsr/ops/dcn/src'


def build_c(force=False):
    src = os.path.join(HERE, 'dcn_ref.c')
    out = os.path.join(HERE, 'libdcn_ref.so')
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = ['gcc', '-O3', '-march=x86-64-v2', '-fopenmp', '-shared', '-fPIC', '-o', out, src, '-lm']
    subprocess.check_call(cmd)
    return out


def build_ref(force=False):
    """Reference DCN ext (deform_conv_ext.cpp, deform_conv_cuda.cpp, deform_conv_cuda_kernel.cu)
    -> oracle/_ref/deform_conv_ext_ref.so, compiled for sm_100a.  Returns path or None."""
    out_dir = os.path.join(HERE, '_ref')
    out = os.path.join(out_dir, 'deform_conv_ext_ref.so')
    if os.path.exists(out) and not force:
        return out
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(out_dir, exist_ok=True)
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    os.environ.setdefault('MAX_JOBS', '4')
    from torch.utils.cpp_extension import load
    build_dir = os.path.join(out_dir, 'build')
    os.makedirs(build_dir, exist_ok=True)
    load(name='deform_conv_ext_ref',
         sources=[os.path.join(REF_SRC, f) for f in
                  ('deform_conv_ext.cpp', 'deform_conv_cuda.cpp', 'deform_conv_cuda_kernel.cu')],
         build_directory=build_dir, is_python_module=False, verbose=False)
    built = os.path.join(build_dir, 'deform_conv_ext_ref.so')
    os.replace(built, out)
    return out


if __name__ == '__main__':
    print(build_c(force=True))
    if '--ref' in sys.argv:
        print(build_ref(force=True))
