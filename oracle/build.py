"""Build recipes for the checker (TEST INFRASTRUCTURE ONLY -- never imported by genpc_b200/).

  python oracle/build.py            # libgenpc_oracle.so (+ oracle/_ref/ when /root/reference exists)

* ``build_oracle()``  compiles oracle/genpc_oracle.c (the CPU restatement) into oracle/libgenpc_oracle.so.
* ``build_ref()``     compiles the UNMODIFIED reference CUDA extensions from the sources where they lie
  under /root/reference (loss_functions/Chamfer3D/{chamfer_cuda.cpp,chamfer3D.cu},
  loss_functions/emd/{emd.cpp,emd_cuda.cu}) for sm_100a into oracle/_ref/{chamfer_3D,emd}/.  Outputs only
  go to oracle/_ref/ (git-ignored, NOT gpurun-ignored, so the built .so travels to the GPU box where
  /root/reference does not exist).  No reference source is copied into the repo.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
ORACLE_SO = os.path.join(HERE, "libgenpc_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    src = os.path.join(HERE, "genpc_oracle.c")
    if not force and _newer(ORACLE_SO, [src]):
        return ORACLE_SO
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
           "-o", ORACLE_SO, src, "-lm"]
    subprocess.check_call(cmd)
    return ORACLE_SO


def ref_so_path(name):
    return os.path.join(REF_DIR, name, name + ".so")


def build_ref(force=False):
    """JIT-compile the reference extensions (needs /root/reference; ~40 s each)."""
    if not os.path.isdir(REF):
        return {}
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    out = {}
    specs = {
        "chamfer_3D": [f"{REF}/loss_functions/Chamfer3D/chamfer_cuda.cpp",
                       f"{REF}/loss_functions/Chamfer3D/chamfer3D.cu"],
        "emd": [f"{REF}/loss_functions/emd/emd.cpp", f"{REF}/loss_functions/emd/emd_cuda.cu"],
    }
    for name, srcs in specs.items():
        so = ref_so_path(name)
        if not force and _newer(so, srcs):
            out[name] = so
            continue
        bdir = os.path.join(REF_DIR, name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=srcs, build_directory=bdir, verbose=False,
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
        out[name] = so
    return out


def install_ref_package(force=False):
    """The base contract's one offline install: the UNMODIFIED reference Python package into baseline/_ref
    (git-ignored, travels to the GPU box) -- from a copy under /tmp because /root/reference is read-only and the
    build writes egg-info; --no-deps because the reference's generator dependencies are not in the wheelhouse."""
    import shutil
    import tempfile

    root = os.path.dirname(HERE)
    target = os.path.join(root, "baseline", "_ref")
    if not os.path.isdir(REF):
        return None
    if not force and os.path.isdir(os.path.join(target, "loss_functions")):
        return target
    tmp = tempfile.mkdtemp(prefix="genpc_ref_")
    src = os.path.join(tmp, "reference")
    shutil.copytree(REF, src, ignore=shutil.ignore_patterns("data", "*.ply", ".git"))
    # packaging only: the reference's setup.py uses find_packages(), which skips loss_functions/Chamfer3D and
    # loss_functions/emd because they have no __init__.py (the reference imports them as namespace packages from its
    # source tree).  Empty __init__.py files in the TEMP COPY make the wheel complete; no reference file is edited.
    for sub in ("Chamfer3D", "emd"):
        init = os.path.join(src, "loss_functions", sub, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    os.makedirs(target, exist_ok=True)
    subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                           "--no-deps", "--upgrade", "--find-links", "/opt/wheelhouse", "--target", target, src])
    shutil.rmtree(tmp, ignore_errors=True)
    return target


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
    print(install_ref_package(force="--force" in sys.argv))
