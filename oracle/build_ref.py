"""TEST INFRASTRUCTURE ONLY -- builds the reference's OWN CPU RoIAlign kernel into oracle/_ref/.

Compiles common/lib/roi_pooling/{vision.cpp,cpu/ROIAlign_cpu.cpp} from /root/reference where they lie
(nothing is copied into the repo; a patched translation unit is piped through /tmp).  The one change
needed for torch >= 1.11 is `input.type()` -> `input.scalar_type()` in the AT_DISPATCH macro at
cpu/ROIAlign_cpu.cpp:242; the CUDA files cannot build at all (THC headers are gone),
so the extension is built CPU-only (no WITH_CUDA).  Output: oracle/_ref/C_ROIPooling_ref*.so
(git-ignored, travels with gpurun snapshots).
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("VLBERT_REFERENCE_ROOT", "/root/reference")
SRC = os.path.join(REF, "common", "lib", "roi_pooling")


def so_path():
    return os.path.join(OUT, "C_ROIPooling_ref" + sysconfig.get_config_var("EXT_SUFFIX"))


def build():
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources not present at %s" % SRC)
    import torch
    from torch.utils import cpp_extension as ce
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="vlb_ref_")
    try:
        # patched translation units live only in /tmp
        for rel in ("vision.cpp", "cpu/ROIAlign_cpu.cpp"):
            s = open(os.path.join(SRC, rel)).read()
            s = s.replace("AT_DISPATCH_FLOATING_TYPES(input.type()", "AT_DISPATCH_FLOATING_TYPES(input.scalar_type()")
            s = s.replace("PYBIND11_MODULE(TORCH_EXTENSION_NAME", "PYBIND11_MODULE(C_ROIPooling_ref")
            dst = os.path.join(tmp, rel.replace("/", "_"))
            open(dst, "w").write(s)
        incs = ["-I" + p for p in ce.include_paths()] + ["-I" + sysconfig.get_paths()["include"], "-I" + SRC,
                                                          "-I" + os.path.join(SRC, "cpu")]
        objs = []
        for f in ("vision.cpp", "cpu_ROIAlign_cpu.cpp"):
            o = os.path.join(tmp, f + ".o")
            cmd = ["g++", "-O2", "-fPIC", "-std=c++17", "-DTORCH_API_INCLUDE_EXTENSION_H",
                   "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI), *incs, "-c",
                   os.path.join(tmp, f), "-o", o]
            subprocess.check_call(cmd)
            objs.append(o)
        libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
        cmd = ["g++", "-shared", "-o", so_path(), *objs, "-L" + libdir, "-lc10", "-ltorch", "-ltorch_cpu",
               "-ltorch_python", "-Wl,-rpath," + libdir]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return so_path()


def load():
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("C_ROIPooling_ref", build())
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


if __name__ == "__main__":
    print(build())
