"""Build hero_b200/libhero_b200_ablate.so: the product objects with csrc/stack.cu recompiled
under -DHERO_ABLATE (env HERO_ABLATE=<mask> then skips kernel families inside the layer stack;
see csrc/stack.cu). Timing experiments only - results of an ablated run are garbage."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hero_b200 import build as hb  # noqa: E402


def main():
    hb.build()
    nvcc = hb._nvcc()
    obj = os.path.join(hb.BUILD_DIR, "stack_ablate.o")
    subprocess.run([nvcc] + hb.NVCC_FLAGS + ["-DHERO_ABLATE", "-c",
                                            os.path.join(hb.CSRC, "stack.cu"), "-o", obj], check=True)
    objs = [os.path.join(hb.BUILD_DIR, s[:-3] + ".o") for s in hb._sources() if s != "stack.cu"]
    out = os.path.join(hb.PKG_DIR, "libhero_b200_ablate.so")
    subprocess.run([nvcc, "-shared", "-o", out, obj] + objs, check=True)
    print("built", out)


if __name__ == "__main__":
    main()
