"""Build libmrefsr_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m mrefsr_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so stays in-tree (mrefsr_b200/lib/) so that it travels
to the GPU box with the repo snapshot; it is git-ignored.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libmrefsr_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
         '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'mrefsr_b200.h')]
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    objdir = os.path.join(LIBDIR, 'obj')
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [NVCC] + ARCH + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        log = r.stdout + r.stderr
        with open(os.path.join(objdir, src + '.ptxas.log'), 'w') as f:
            f.write(log)
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs + ['-Xcompiler', '-fPIC', '-cudart', 'static']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


TORCH_EXT_DIR = os.path.join(HERE, 'torch_ext')
TORCH_EXT = os.path.join(TORCH_EXT_DIR, 'deform_conv_ext.so')


def build_torch_ext(force=False, verbose=False):
    """The torch extension module `deform_conv_ext` (csrc/torch_ext/deform_conv_ext.cpp): the reference's five pybind
    exports on top of libmrefsr_b200.so -- the file that replaces the reference's own deform_conv_ext build.  Built
    in-tree (mrefsr_b200/torch_ext/deform_conv_ext.so, git-ignored) so that it travels to the GPU box; it finds the
    kernel library through an $ORIGIN-relative rpath."""
    src = os.path.join(CSRC, 'torch_ext', 'deform_conv_ext.cpp')
    build(force=False)
    if (not force and os.path.exists(TORCH_EXT) and os.path.getmtime(TORCH_EXT) >= os.path.getmtime(src) and
            os.path.getmtime(TORCH_EXT) >= os.path.getmtime(os.path.join(HERE, '..', 'include', 'mrefsr_b200.h'))):
        return TORCH_EXT
    os.makedirs(TORCH_EXT_DIR, exist_ok=True)
    bdir = os.path.join(TORCH_EXT_DIR, 'build')
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault('MAX_JOBS', '4')
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    from torch.utils.cpp_extension import load
    load(name='deform_conv_ext', sources=[src], build_directory=bdir, is_python_module=False, verbose=verbose,
         with_cuda=True, extra_cflags=['-O2'],
         extra_ldflags=['-L' + LIBDIR, '-lmrefsr_b200', # ninja ($$) and the shell (quotes) both see this string before the linker does
                        "-Wl,-rpath,'$$ORIGIN/../lib'", "-Wl,-rpath,'$$ORIGIN/../../lib'", '-Wl,--no-as-needed'])
    os.replace(os.path.join(bdir, 'deform_conv_ext.so'), TORCH_EXT)
    return TORCH_EXT


def load_torch_ext():
    """Import the built extension module (raises with the build command if it is missing: no fallback)."""
    import importlib.util
    if not os.path.exists(TORCH_EXT):
        raise RuntimeError('mrefsr_b200: %s is missing -- build it with `python -m mrefsr_b200.build --torch-ext`' % TORCH_EXT)
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location('deform_conv_ext', TORCH_EXT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
    if '--torch-ext' in sys.argv:
        print(build_torch_ext(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
