"""Build libwiflow_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SOURCES = ['wf_conv.cu', 'wf_group.cu', 'wf_slide.cu', 'wf_slabtc.cu', 'wf_tc.cu', 'wf_thin.cu', 'wf_elem.cu', 'wf_attn.cu', 'wf_data.cu', 'wf_model.cu']
LIB = os.path.join(HERE, 'libwiflow_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'wiflow_b200.h'))
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace('.cu', '.o'))
        objs.append(o)
        if force or any(_newer(d, o) for d in headers + [s]):
            cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError(f'nvcc failed on {failed}')
    manifest = os.path.join(CSRC, '.link_manifest')          # relink when the list of objects changes, not only when one is rebuilt
    want = ' '.join(SOURCES)
    have = open(manifest).read() if os.path.exists(manifest) else ''
    if force or procs or want != have or not os.path.exists(LIB) or any(_newer(o, LIB) for o in objs):
        subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-Xcompiler', '-fPIC', '-lcudart'])
        with open(manifest, 'w') as f:
            f.write(want)
    # native self-test of the tcgen05 kernels (tests/test_gpu_native.py runs it on the GPU box)
    st_src = os.path.join(HERE, '..', 'tests', 'native', 'tc_selftest.cu')
    st_exe = os.path.join(HERE, '..', 'tests', 'native', 'tc_selftest')
    tc_obj = os.path.join(CSRC, 'wf_tc.o')
    if os.path.exists(st_src) and (force or procs or any(_newer(d, st_exe) for d in [st_src, tc_obj])):
        subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O2', '-std=c++17', '-Wno-deprecated-gpu-targets',
                               '-o', st_exe, st_src, tc_obj, '-lcudart'])
    sl_src = os.path.join(HERE, '..', 'tests', 'native', 'slab_selftest.cu')
    sl_exe = os.path.join(HERE, '..', 'tests', 'native', 'slab_selftest')
    sl_obj = os.path.join(CSRC, 'wf_slabtc.o')
    if os.path.exists(sl_src) and (force or procs or any(_newer(d, sl_exe) for d in [sl_src, sl_obj])):
        subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O2', '-std=c++17', '-Wno-deprecated-gpu-targets',
                               '-o', sl_exe, sl_src, sl_obj, '-lcudart'])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
