"""Build every native artefact IN-TREE (so that the .so files travel to the GPU box):

  pdmp3_b200/libpdmp3_b200.so   the product: plain-C host side + sm_100a kernels (nvcc, sm_100a only)
  tools/libp3synth.so           synthetic stream generator (test/bench infrastructure)
  tools/tc_trial/tc_matrix      the tcgen05 trial of the matrixing stage (measurement tool)
  oracle/libp3_oracle.so        CPU restatement of the reference (checker)
  oracle/_ref/*                 the compiled reference + tap harness (only where /root/reference exists)
"""
import os, subprocess, sys, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "pdmp3_b200", "csrc")
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd[:3]))
    return r.stdout


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_product(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    gcc = shutil.which("gcc") or "gcc"
    out = os.path.join(ROOT, "pdmp3_b200", "libpdmp3_b200.so")
    srcs_c = ["p3_tables.c", "p3_parse.c", "p3_api.c"]
    srcs_cu = ["p3_kernels.cu", "p3_fused.cu", "p3_hop.cu", "p3_cabi.cu"]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if not force and not _newer(out, deps):
        return out
    bdir = os.path.join(ROOT, "build"); os.makedirs(bdir, exist_ok=True)
    objs = []
    for s in srcs_c:
        o = os.path.join(bdir, s + ".o")
        _run([gcc, "-O2", "-Wall", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, s), "-o", o])
        objs.append(o)
    for s in srcs_cu:
        o = os.path.join(bdir, s + ".o")
        log = _run([nvcc] + NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "550", "-Xptxas", "-v", "-Xcompiler", "-fPIC",
                    "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, s), "-o", o])
        if verbose:
            print(log)
        objs.append(o)
    _run([nvcc] + NVCC_ARCH + ["-shared", "-o", out] + objs + ["-lpthread", "-lm", "-ldl"])
    return out


def build_tools(force=False):
    gcc = shutil.which("gcc") or "gcc"
    out = os.path.join(ROOT, "tools", "libp3synth.so")
    deps = [os.path.join(ROOT, "tools", "p3_synth.c"), os.path.join(ROOT, "tools", "p3_synth.h"), os.path.join(CSRC, "p3_tables.c"), os.path.join(CSRC, "p3_huffcodes.inc")]
    if force or _newer(out, deps):
        _run([gcc, "-O2", "-Wall", "-fPIC", "-shared", "-o", out, deps[0], deps[2], "-lm", "-lpthread"])
    return out


def build_trials(force=False):
    """tools/tc_trial/tc_matrix: the stand-alone tcgen05 trial of the matrixing stage (measurement tool, not product)"""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = os.path.join(ROOT, "tools", "tc_trial", "tc_matrix.cu"); out = os.path.join(ROOT, "tools", "tc_trial", "tc_matrix")
    if force or _newer(out, [src, os.path.join(CSRC, "p3_xform.cuh"), os.path.join(CSRC, "p3_lee.inc")]):
        _run([nvcc] + NVCC_ARCH + ["-O3", "-lineinfo", "-std=c++17", "-diag-suppress", "550", "-I", CSRC, "-o", out, src])
    return out


def build_oracle():
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_all(force=False, verbose=False):
    build_product(force, verbose)
    build_tools(force)
    build_trials(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built")
