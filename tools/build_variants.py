"""A/B builds of the hot kernel: recompiles only the listed translation units
with extra -D switches and links them with the base objects.

usage: python tools/build_variants.py [--tu k_ung_psi1.cu] name1="-DX=1 -DY=2" name2=""
Writes build_variants/<name>.so (+ .log with ptxas -v); run with BSSM_B200_LIB=<path>.
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bssm_b200 import build as B

def main():
    args = sys.argv[1:]
    tus = ["k_ung_psi1.cu"]
    if args and args[0] == "--tu":
        tus = args[1].split(","); args = args[2:]
    B.build()
    out = os.path.join(B.ROOT, "build_variants"); os.makedirs(out, exist_ok=True)
    def one(spec):
        name, _, defs = spec.partition("=")
        objs = []
        log = ""
        for src in B.SOURCES:
            if src in tus:
                obj = os.path.join(out, name + "_" + src[:-3] + ".o")
                res = subprocess.run(B.compile_cmd(src, obj, True, defs.split()), capture_output=True, text=True)
                log += res.stdout + res.stderr
                if res.returncode: return name, res.returncode
                objs.append(obj)
            else:
                objs.append(B._obj_of(src))
        open(os.path.join(out, name + ".log"), "w").write(log)
        res = subprocess.run(B.link_cmd(objs, os.path.join(out, name + ".so")), capture_output=True, text=True)
        return name, res.returncode
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        for name, rc in ex.map(one, args): print(name, "rc", rc)

if __name__ == "__main__":
    main()
