"""Builds the reference's own deformable-conv CUDA extension (UNMODIFIED sources, compiled where they lie under
/root/reference) for sm_100a into oracle/_ref/ -- the real reference op as a GPU cross-check of the oracle and of
`cdn_deform_conv_forward_f32` (tests/test_gpu_reference_ext.py), and as the reference arm of the isolated-layer timing.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.build() and bench tooling may call this; the product never imports
anything under oracle/.  The reference's build system (lib/models/external/Makefile + build_dcn.py, torch 1.x ffi) is
not run: the two translation units are compiled directly,
    lib/models/external/src/dcn_deform_conv_cuda.cpp          (host glue, pybind module `dcn_deform_conv_cuda`)
    lib/models/external/src/dcn_deform_conv_cuda_kernel.cu    (through oracle/ref_dcn_kernel_wrapper.cu)
with -DAT_CHECK=TORCH_CHECK -DIntList=IntArrayRef (API renames since torch 1.x).  /root/reference exists only in the build
container; the built module travels to the GPU box with the repo snapshot (oracle/_ref is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/lib/models/external/src"
NAME = "dcn_deform_conv_cuda"


def module_path():
    return os.path.join(OUT, NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def available():
    return os.path.exists(module_path())


def build(force=False, verbose=False):
    """Returns the module path, or None when the reference tree is absent (GPU box: use the prebuilt file)."""
    so = module_path()
    cpp, cu = os.path.join(REF_SRC, NAME + ".cpp"), os.path.join(REF_SRC, NAME + "_kernel.cu")
    if not (os.path.exists(cpp) and os.path.exists(cu)):
        return so if os.path.exists(so) else None
    wrapper = os.path.join(HERE, "ref_dcn_kernel_wrapper.cu")
    if not force and os.path.exists(so) and os.path.getmtime(so) > max(os.path.getmtime(p) for p in (cpp, cu, wrapper, __file__)):
        return so
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    inc = sum((["-I", p] for p in ce.include_paths(device_type="cuda") + [sysconfig.get_paths()["include"]]), [])
    defs = ["-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H", "-DAT_CHECK=TORCH_CHECK", "-DIntList=IntArrayRef",
            "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    o_cpp, o_cu = os.path.join(OUT, "host.o"), os.path.join(OUT, "kernel.o")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmds = [
        ["g++", "-std=c++17", "-O2", "-fPIC", "-w", "-c", cpp, "-o", o_cpp] + defs + inc,
        [nvcc, "-std=c++17", "-O2", "-c", wrapper, "-o", o_cu, "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
         "-w", "-Xcompiler", "-fPIC", '-DREF_DCN_KERNEL_CU="%s"' % cu] + defs + inc,
        ["g++", "-shared", o_cpp, o_cu, "-o", so] + sum((["-L", p] for p in ce.library_paths(device_type="cuda")), []) +
        ["-lc10", "-lc10_cuda", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-ltorch_python", "-lcudart"],
    ]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("building the reference extension failed (%s)" % os.path.basename(cmd[-1]))
    for o in (o_cpp, o_cu):
        os.remove(o)
    return so


REF_ROOT = "/root/reference"
PY_OUT = os.path.join(OUT, "py")
# what oracle/ref_harness.py imports (and their package __init__ files): the network, the operator modules, the
# quantizer, decode and the post-processing helpers.  Nothing else of the tree is needed to run the path.
PY_FILES = [
    "lib/models/__init__.py", "lib/models/decode.py", "lib/models/utils.py",
    "lib/models/networks/__init__.py", "lib/models/networks/shufflenetv2_dcn.py",
    "lib/models/external/__init__.py", "lib/models/external/modules/__init__.py",
    "lib/models/external/modules/dcn_deform_conv.py", "lib/models/external/functions/__init__.py",
    "lib/models/external/functions/dcn_deform_conv.py",
    "lib/utils/__init__.py", "lib/utils/post_process.py", "lib/utils/image.py", "lib/utils/ddd_utils.py",
    "portable_quantizer/__init__.py", "portable_quantizer/quant_modules.py",
    "portable_quantizer/quantization_utils/__init__.py", "portable_quantizer/quantization_utils/quant_utils.py",
    "portable_quantizer/quantization_utils/quantize_model.py",
]


def py_root():
    """Directory holding an importable reference tree: /root/reference in the build container, the installed copy
    oracle/_ref/py on the GPU box (None when neither exists)."""
    if os.path.isdir(os.path.join(REF_ROOT, "lib", "models")):
        return REF_ROOT
    if os.path.isdir(os.path.join(PY_OUT, "lib", "models")):
        return PY_OUT
    return None


def build_py(force=False):
    """Installs the reference's Python files of the path, UNMODIFIED, into the git-ignored oracle/_ref/py so that
    `bench.py --impl reference` / `cpu_baseline` can time the reference's own PyTorch CPU forward on the GPU box's host
    cores (the counterpart of `pip install --target baseline/_ref`: the reference has no setup.py).  Build container
    only; the copy is never committed (oracle/_ref is git-ignored) and never imported by the product."""
    import shutil
    if not os.path.isdir(os.path.join(REF_ROOT, "lib", "models")):
        return PY_OUT if os.path.isdir(PY_OUT) else None
    for rel in PY_FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(PY_OUT, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):             # namespace-style directories of the reference: keep them importable
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                continue
            raise FileNotFoundError(src)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return PY_OUT


NMS_PYX = "/root/reference/lib/models/external/nms.pyx"


def build_nms(force=False):
    """The reference's Cython NMS (lib/models/external/nms.pyx, unmodified): cython -> C in oracle/_ref, gcc -> module
    `nms`.  Used by oracle/make_golden.py to generate tests/golden/soft_nms_kat.npz; never shipped or imported by the product."""
    import numpy
    so = os.path.join(OUT, "nms" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(NMS_PYX):
        return so if os.path.exists(so) else None
    if not force and os.path.exists(so) and os.path.getmtime(so) > os.path.getmtime(NMS_PYX):
        return so
    os.makedirs(OUT, exist_ok=True)
    c_file = os.path.join(OUT, "nms.c")
    # numpy 2 dropped the `np.int_t` ctypedef that the (unused here) hard-NMS function of the file names at :32; the file
    # is cythonised from a temporary copy with that one identifier renamed (deleted again below, never committed) --
    # the counterpart of the two -D renames of the CUDA extension.  soft_nms (:77-170) is compiled as written.
    tmp_pyx = os.path.join(OUT, "nms.pyx")
    with open(NMS_PYX) as f:
        src = f.read().replace("np.int_t", "np.intp_t")
    with open(tmp_pyx, "w") as f:
        f.write(src)
    cmds = [[sys.executable, "-m", "cython", "-3", tmp_pyx, "-o", c_file],
            ["gcc", "-O2", "-fPIC", "-shared", "-w", c_file, "-o", so, "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include()]]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("building the reference nms.pyx failed")
    os.remove(c_file)
    os.remove(tmp_pyx)
    return so


def load_nms():
    import importlib.util
    so = os.path.join(OUT, "nms" + sysconfig.get_config_var("EXT_SUFFIX"))
    spec = importlib.util.spec_from_file_location("nms", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Imports the built module (needs torch imported first).  Raises if it was never built."""
    import importlib.util
    import torch  # noqa: F401  (the module links against libtorch)
    so = module_path()
    if not os.path.exists(so):
        raise FileNotFoundError(so + " (run oracle/build_ref.py in the build container)")
    spec = importlib.util.spec_from_file_location(NAME, so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
