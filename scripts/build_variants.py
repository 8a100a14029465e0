"""Build tuning variants of libsphb200.so HERE (nvcc cross-compiles; no GPU time spent on compiling) into
spheral_b200/variants/libsphb200_<tag>.so; scripts/gpu_ab.sh benches each through SPHB200_LIB.

    python scripts/build_variants.py tag1="-DX=1 -DY=2" tag2="-DX=3" ...
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spheral_b200 import build as B  # noqa: E402

out = os.path.join(ROOT, "spheral_b200", "variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    tag, flags = spec.split("=", 1)
    objdir = os.path.join(out, "obj_" + tag)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in B.SOURCES:
        obj = os.path.join(objdir, src + ".o")
        procs.append((src, obj, subprocess.Popen([B.NVCC] + B.FLAGS + flags.split() + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", obj],
                                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        o, _ = p.communicate()
        if p.returncode:
            sys.exit("nvcc failed for %s [%s]\n%s" % (src, tag, o))
        open(os.path.join(objdir, src + ".log"), "w").write(o)
        objs.append(obj)
    lib = os.path.join(out, "libsphb200_%s.so" % tag)
    subprocess.check_call([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                                                                  "-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    shutil.rmtree(objdir + "_keep", ignore_errors=True)
    print(lib)
