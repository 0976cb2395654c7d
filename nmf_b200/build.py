"""Builds libnmf_b200.so (sm_100a) in-tree:  python -m nmf_b200.build [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnmf_b200.so")
SOURCES = ["nmf_kernels.cu", "nmf_train.cu", "nmf_env_bwd.cu", "nmf_normals_bwd.cu", "nmf_shade_bwd.cu", "nmf_mf_train.cu", "nmf_bench.cu", "nmf_repack.cu"]
DEPS = ["nmf_bench.cu", "nmf_repack.cu", "nmf_kernels.cu", "nmf_train.cu", "nmf_env_bwd.cu", "nmf_normals_bwd.cu", "nmf_shade_bwd.cu", "nmf_mf_train.cu", "nmf_render_ws.cuh", "nmf_microfacet_bwd.cuh", "nmf_train.cuh", "nmf_mlp_tc.cuh", "nmf_mlp_tc_bwd.cuh", "nmf_math.cuh", "nmf_field.cuh", os.path.join("..", "..", "include", "nmf_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--threads", "6"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    extra = os.environ.get("NMF_NVCC_EXTRA", "").split()      # experiments only (e.g. -DNMF_SHADE_UNROLL=2)
    cmd = [nvcc()] + NVCC_FLAGS + extra + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", OUT)
