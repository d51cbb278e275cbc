"""
Builds libirsgmcmc.so (hand-written CUDA for sm_100a, C ABI in include/irsgmcmc.h) in-tree with nvcc.
nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repository snapshot.
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libirsgmcmc.so')

NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC',
              '-diag-suppress', '177']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.isfile(nvcc):
        raise RuntimeError('nvcc not found: libirsgmcmc.so cannot be built (there is no non-CUDA build)')
    return nvcc


def _stale(target, sources):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force=False, verbose=False):
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    deps = sources + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(ROOT, 'include', 'irsgmcmc.h')]
    if not force and not _stale(LIB, deps):
        return LIB
    cmd = [_nvcc(), '-shared'] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + sources + ['-o', LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_host_emulation(force=False):
    """test infrastructure: the library's __host__ __device__ arithmetic in CPU loops (tests/host_emul)"""
    src = os.path.join(ROOT, 'tests', 'host_emul', 'host_emul.cu')
    out = os.path.join(ROOT, 'tests', 'host_emul', 'libirs_host_emul.so')
    deps = [src] + glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(ROOT, 'include', 'irsgmcmc.h')]
    if not force and not _stale(out, deps):
        return out
    cmd = [_nvcc(), '-shared', '-O2', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-Xcompiler', '-fPIC',
           '-diag-suppress', '20013', src, '-o', out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc (host emulation) failed:\n' + res.stdout + res.stderr)
    return out


def build_oracle_c(force=False):
    """test infrastructure: the plain-C restatement of the bit-exact nearest-neighbour warp (oracle/nearest_warp.c)"""
    src = os.path.join(ROOT, 'oracle', 'nearest_warp.c')
    out = os.path.join(ROOT, 'oracle', 'libnearest_warp.so')
    if not os.path.isfile(src):
        return None
    if not force and not _stale(out, [src]):
        return out
    res = subprocess.run(['gcc', '-O2', '-shared', '-fPIC', '-ffp-contract=off', src, '-o', out, '-lm'],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('gcc (oracle) failed:\n' + res.stdout + res.stderr)
    return out


if __name__ == '__main__':
    print(build_library(force=True, verbose=False))
    print(build_host_emulation(force=True))
    print(build_oracle_c(force=True))
