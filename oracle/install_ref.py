"""TEST / BASELINE INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference under the git-ignored `baseline/_ref/`.

`python -m pip install --target baseline/_ref /root/reference` is not possible (the reference has neither setup.py nor
pyproject.toml at its root: "Directory '/root/reference' is not installable"), so this recipe does what that install would:

  1. copies the reference's Python packages `models/`, `ops/`, `utils/` and `configuration/` byte for byte into
     `baseline/_ref/` (git-ignored, NOT gpurun-ignored: it travels to the GPU box; nothing of it enters the history);
  2. builds the reference's own CUDA extension `MultiScaleDeformableAttention` (ops/src/**: vision.cpp,
     cpu/ms_deform_attn_cpu.cpp, cuda/ms_deform_attn_cuda_t.cu) for sm_100a into `baseline/_ref/` -- from a SHIMMED COPY of
     ops/src under `baseline/_ref/_msda_build/` whose only change is `value.type()` -> `value.scalar_type()` inside the two
     AT_DISPATCH_FLOATING_TYPES_AND_HALF lines (ops/src/cuda/ms_deform_attn_cuda_t.cu:64,134): torch >= 2.x no longer converts
     DeprecatedTypeProperties to ScalarType there.  The kernels are untouched.

Run in the build container (where /root/reference is mounted):  python oracle/install_ref.py [--no-ext]
Consumers: oracle/ref_import.py (reference decoder for `bench.py --impl reference` and the fixture generators) and
tests/test_msda_ref_gpu.py (reference MSDA kernel = forward oracle + kernel to beat).  The product never imports it.
"""
import argparse
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CQVAD_REF_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
SHIM = (("AT_DISPATCH_FLOATING_TYPES_AND_HALF(value.type(),", "AT_DISPATCH_FLOATING_TYPES_AND_HALF(value.scalar_type(),"),)


def _writable(path):
    for r, ds, fs in os.walk(path):
        for n in ds + fs:
            q = os.path.join(r, n)
            os.chmod(q, os.stat(q).st_mode | 0o200)
    os.chmod(path, os.stat(path).st_mode | 0o200)


def copy_python():
    os.makedirs(DST, exist_ok=True)
    _writable(DST)                      # the reference tree is mounted read-only; copytree preserves the modes
    for pkg in ("models", "ops", "utils", "configuration"):
        d = os.path.join(DST, pkg)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(SRC, pkg), d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "build"))
    for f in ("LICENSE", "NOTICE", "README.md"):
        shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
    _writable(DST)


def build_msda_ext():
    """nvcc cross-compiles for sm_100a without a GPU; torch.utils.cpp_extension supplies the include / link flags."""
    bdir = os.path.join(DST, "_msda_build")
    if os.path.isdir(bdir):
        shutil.rmtree(bdir)
    shutil.copytree(os.path.join(SRC, "ops", "src"), os.path.join(bdir, "src"))
    _writable(bdir)
    cu = os.path.join(bdir, "src", "cuda", "ms_deform_attn_cuda_t.cu")
    text = open(cu).read()
    n = 0
    for a, b in SHIM:
        n += text.count(a)
        text = text.replace(a, b)
    assert n == 2, f"expected 2 shim sites, found {n}"
    open(cu, "w").write(text)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", str(os.cpu_count() or 4))
    from torch.utils.cpp_extension import load
    src = os.path.join(bdir, "src")
    sources = glob.glob(os.path.join(src, "*.cpp")) + glob.glob(os.path.join(src, "cpu", "*.cpp")) + glob.glob(os.path.join(src, "cuda", "*.cu"))
    load(name="MultiScaleDeformableAttention", sources=sources, extra_include_paths=[src],
         extra_cflags=["-DWITH_CUDA"],
         extra_cuda_cflags=["-DWITH_CUDA", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                            "-D__CUDA_NO_HALF2_OPERATORS__", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=bdir, is_python_module=False, verbose=False)
    so = os.path.join(bdir, "MultiScaleDeformableAttention.so")
    assert os.path.exists(so), "extension build produced no .so"
    shutil.copy2(so, os.path.join(DST, "MultiScaleDeformableAttention.so"))
    for junk in glob.glob(os.path.join(bdir, "*.o")):
        os.remove(junk)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-ext", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(SRC):
        print(f"install_ref: {SRC} not present (GPU box?) -- using the prebuilt baseline/_ref as is")
        return 0
    copy_python()
    if not a.no_ext:
        build_msda_ext()
    subprocess.call(["ls", "-la", DST])
    return 0


if __name__ == "__main__":
    sys.exit(main())
