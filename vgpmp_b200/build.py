"""Build libvgpmp_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

Every translation unit is compiled to its own object (in parallel, only when it or a header is newer than the object) and
the objects are linked into vgpmp_b200/lib/libvgpmp_b200.so.  Objects and the library are git-ignored; the library
travels to the GPU box with the tree."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = [PKG / "csrc" / n for n in ("cabi.cu", "kinematics.cu", "gp.cu", "sampler_tc.cu", "mesh_sdf.cu")]
HDR = [PKG / "csrc" / "common.cuh", PKG / "csrc" / "device_utils.cuh", PKG.parent / "include" / "vgpmp_b200.h"]
LIB = PKG / "lib" / "libvgpmp_b200.so"
OBJ = PKG / "lib" / "obj"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--fmad=true", "-Xptxas", "-v"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libvgpmp_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(p.stat().st_mtime > t for p in deps)


def needs_build() -> bool:
    return _stale(LIB, SRC + HDR)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    OBJ.mkdir(parents=True, exist_ok=True)
    nvcc = nvcc_path()
    logs = {}

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        if not force and not _stale(obj, [src] + HDR):
            return obj, None
        cmd = [nvcc, *NVCC_FLAGS, "-c", "-o", str(obj), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs[src.name] = " ".join(cmd) + "\n" + res.stdout + res.stderr
        return obj, (res if res.returncode != 0 else None)

    with ThreadPoolExecutor(max_workers=len(SRC)) as pool:
        results = list(pool.map(compile_one, SRC))
    log_text = "".join(f"==== {k}\n{v}" for k, v in sorted(logs.items()))
    failed = [r for _, r in results if r is not None]
    if failed:
        (PKG / "lib" / "build.log").write_text(log_text)
        raise RuntimeError("nvcc failed:\n" + "\n".join(r.stderr[-4000:] for r in failed))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *[str(o) for o, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (PKG / "lib" / "build.log").write_text(log_text + "==== link\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stderr[-4000:])
    if verbose:
        print(log_text)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
