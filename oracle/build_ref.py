"""CPU/GPU ORACLE BUILD RECIPE (test infrastructure only; never on the product path).

Compiles the reference's ONLY native component, ``csrc/csrc/cuda_api.cu`` (the pybind11 torch
extension ``tiny_api_cuda`` with ``update_flatten_view``), UNMODIFIED and from where it lies under
``/root/reference``, for sm_100a, into ``oracle/_ref/tiny_api_cuda.so`` (git-ignored, shipped to
the GPU box by gpurun like our own ``.so``).  The reference's own build system (``csrc/build.py``,
setuptools + ninja, ``-gencode arch=compute_80,code=sm_80``) is not run: this is one direct nvcc
command with torch's include/library paths and the reference's own nvcc flags, only the arch changed.

Second artefact: ``oracle/_ref/ref_pyramidkv_utils.py``, a byte copy of the reference's policy file
(git-ignored like the ``.so``: no reference source enters the history).  It is imported ONLY by
``bench.py --impl reference`` / ``cpu_baseline`` (the timed CPU baseline, ``kind: "reference"``) and by
the parity tests as a checker next to ``oracle/kvc_oracle.py``; the product never imports it.

    python oracle/build_ref.py            # no-op when /root/reference is absent (the GPU box)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KVC_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(ROOT, "oracle", "_ref")


def build_tiny_api_cuda(force: bool = False) -> str | None:
    src = os.path.join(REF, "csrc", "csrc", "cuda_api.cu")
    out = os.path.join(OUT_DIR, "tiny_api_cuda.so")
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames \
        else ce.include_paths(cuda=True)
    libdirs = ce.library_paths(device_type="cuda") if "device_type" in ce.library_paths.__code__.co_varnames \
        else ce.library_paths(cuda=True)
    import torch
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           # the reference's own nvcc flags (csrc/build.py:48-62)
           "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
           "-U__CUDA_NO_BFLOAT16_CONVERSIONS__", "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
           "-lineinfo", "-DTORCH_EXTENSION_NAME=tiny_api_cuda", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-I", os.path.join(REF, "csrc", "csrc"), "-I", os.path.join(REF, "csrc", "include"),
           "-I", sysconfig.get_paths()["include"]]
    for p in inc:
        cmd += ["-I", p]
    for p in libdirs:
        cmd += ["-L", p, "-Xlinker", f"-rpath={p}"]
    cmd += [src, "-o", out, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building the reference's cuda_api.cu failed:\n" + proc.stdout[-4000:])
    return out


def fetch_policy_file() -> str | None:
    """Byte copy of the reference's policy file into oracle/_ref (git-ignored): the timed CPU baseline."""
    src = os.path.join(REF, "pyramidkv", "pyramidkv_utils.py")
    out = os.path.join(OUT_DIR, "ref_pyramidkv_utils.py")
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    if not os.path.exists(out) or open(src, "rb").read() != open(out, "rb").read():
        shutil.copyfile(src, out)
    return out


def main() -> None:
    print("policy file:", fetch_policy_file())
    print("tiny_api_cuda:", build_tiny_api_cuda(force="--force" in sys.argv))


if __name__ == "__main__":
    main()
