"""Builds libyolo_b200.so (hand-written CUDA for sm_100a + the C-ABI of include/yolo_b200.h) in-tree with nvcc.

    python -m yolo_b200.build        (or: __graft_entry__.build())

The .so sits next to this file so that it travels to the GPU box with the repository snapshot.
nvcc cross-compiles without a GPU.  `-gencode arch=compute_100a,code=sm_100a` is spelled out: plain
`-arch=sm_100a` injects an extra compute_100 target in this image and the tcgen05 code does not build for it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libyolo_b200.so")
SOURCES = ["yolo_b200.cu", "conv_direct.cu", "conv_first.cu", "conv_fs.cu", "conv_umma.cu", "conv_ws.cu", "conv_rp.cu", "conv_wsp.cu", "conv_ws2.cu", "conv_ws3.cu", "quantize.cu", "resize.cu", "head.cu", "graph.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"]


def nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "yolo_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        cmd = [nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (s, out))
        if p.returncode:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-cudart", "static"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
