"""Builds libdeepxi_b200.so (hand-written sm_100a CUDA, C ABI of include/deepxi_b200.h) in-tree with nvcc.

    python -m deepxi_b200.build            # incremental
    python -m deepxi_b200.build --force

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# DXI_DEBUG_BUILD=1: the tuning build (-DDXI_ENABLE_DEBUG: phase clocks, per-CTA timelines, the dxi_debug_* entry points of
# include/deepxi_b200_debug.h) as libdeepxi_b200_dbg.so next to the product library; DXI_LIB selects it at run time.
DEBUG = bool(os.environ.get('DXI_DEBUG_BUILD'))
OBJ = os.path.join(HERE, 'build_dbg' if DEBUG else 'build')
LIB = os.path.join(HERE, 'libdeepxi_b200_dbg.so' if DEBUG else 'libdeepxi_b200.so')
SOURCES = ['common.cu', 'gain.cu', 'stft.cu', 'net.cu', 'tcn_f32.cu', 'tcn_umma.cu', 'tcn_chain.cu', 'umma_selftest.cu', 'mhanet.cu', 'mha_umma.cu', 'attn_umma.cu', 'train_tgt.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr'] + (['-DDXI_ENABLE_DEBUG'] if DEBUG else [])


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'deepxi_b200.h'))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace('.cu', '.o'))
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed on %s' % src)
    objs = [os.path.join(OBJ, s.replace('.cu', '.o')) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
