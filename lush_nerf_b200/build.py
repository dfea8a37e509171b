"""Build liblushnerf_b200.so (sm_100a only) in-tree with nvcc.

    python -m lush_nerf_b200.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/lushnerf_b200.h): no torch headers, static
cudart.  It is loaded with ctypes by lush_nerf_b200._lib.  Built artefacts are git-ignored but travel
to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "liblushnerf_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
if os.environ.get("LUSH_CHAIN_TRACE", "0") not in ("", "0"):
    COMMON.append("-DLUSH_CHAIN_TRACE")      # timeline probes of the chain kernels (tools/trace_chain.py); use --force
# experiment builds (tools/jobs/*): LUSH_BUILD_DEFINES="-DLUSH_EXP_NO_STASH ..." LUSH_BUILD_OUT=lib_exp.so python -m lush_nerf_b200.build
if os.environ.get("LUSH_BUILD_DEFINES"):
    COMMON += os.environ["LUSH_BUILD_DEFINES"].split()
if os.environ.get("LUSH_BUILD_OUT"):
    LIB = os.path.abspath(os.environ["LUSH_BUILD_OUT"])
    OBJ = OBJ + "_" + os.path.splitext(os.path.basename(LIB))[0]
# per-file extra flags: the scan / sampler / ray kernels follow the reference's unfused fp32 op
# sequence (bit-exact sample_pdf indices), so FMA contraction is disabled there.
EXTRA = {
    "rays.cu": ["--fmad=false"],
    "composite.cu": ["--fmad=false"],
    "sampler.cu": ["--fmad=false"],
}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "lushnerf_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    spath = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(spath), _deps_mtime())):
        return obj, False
    cmd = [NVCC, "-c", spath, "-o", obj] + COMMON + EXTRA.get(src, [])
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if verbose or res.stderr.strip():
        sys.stderr.write(res.stderr)
    return obj, True


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(r for _, r in results)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
