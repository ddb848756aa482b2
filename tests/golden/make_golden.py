"""Copies the reference's golden *.check values for the fixtures this repo pins into small
JSON files (run HERE, where /root/reference exists; the GPU box only sees the JSON).

    python tests/golden/make_golden.py
"""
import json
import os

REF = "/root/reference/test/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(path):
    """Real arrays as floats; complex ones (Fortran list-directed "(re,im)") as [re, im] pairs."""
    out = []
    with open(path) as f:
        for x in f.read().split():
            if x.startswith("("):
                re_, im_ = x.strip("()").split(",")
                out.append([float(re_), float(im_)])
            else:
                out.append(float(x))
    return out


def inputs(path):
    out = {}
    with open(path) as f:
        for line in f:
            if "=" not in line:
                continue
            k, v = line.split("!")[0].split("=", 1)
            out[k.strip()] = v.strip()
    return out


def umatrix_lines(path):
    """Operator lines of a umatrix.restart file: [oi, si, oj, sj, ok, sk, ol, sl, U]."""
    out = []
    with open(path) as f:
        for raw in f:
            t = raw.split()
            if len(t) == 9 and t[1] in "ud":
                out.append([int(t[0]), t[1], int(t[2]), t[3], int(t[4]), t[5], int(t[6]), t[7], float(t[8])])
    return out


def twobody_hk_calls(path):
    """Arguments of the ed_add_twobody_operator calls of the driver's set_twobody_hk()."""
    import re
    out = []
    pat = re.compile(r'ed_add_twobody_operator\((\d),"(\w)",(\d),"(\w)",(\d),"(\w)",(\d),"(\w)",([-0-9.d+eE]+)\)')
    with open(path) as f:
        for m in pat.finditer(f.read()):
            a = m.groups()
            out.append([int(a[0]), a[1], int(a[2]), a[3], int(a[4]), a[5], int(a[6]), a[7],
                        float(a[8].replace("d", "e"))])
    return out


def main():
    checks = {"NORMAL_NORMAL": ("evals", "dens", "docc", "energy", "doubles", "imp", "Sigma_momenta"),
              "HYBRID_NONSU2": ("evals", "dens", "docc", "energy", "doubles", "imp", "magX",
                                "Sigma11_momenta", "Sigma12_momenta"),
              "INEQ_NORMAL_NORMAL": ("dens", "docc", "energy", "doubles", "Sigma_momenta"),
              "HYBRID_NORMAL": ("evals", "dens", "docc", "energy", "doubles", "imp", "Sigma_momenta"),
              "REPLICA_NORMAL": ("evals", "dens", "docc", "energy", "doubles", "imp", "Sigma_momenta"),
              "GENERAL_NORMAL": ("evals", "dens", "docc", "energy", "doubles", "imp", "Sigma_momenta"),
              "NORMAL_NONSU2": ("evals", "dens", "docc", "energy", "doubles", "imp", "magX"),
              "NORMAL_SUPERC": ("evals", "dens", "docc", "phisc", "energy", "doubles", "imp"),
              "HYBRID_SUPERC": ("evals", "dens", "docc", "phisc", "energy", "doubles", "imp"),
              "REPLICA_NONSU2": ("evals", "dens", "docc", "energy", "doubles", "imp", "exciton"),
              "GENERAL_NONSU2": ("evals", "dens", "docc", "energy", "doubles", "imp", "exciton"),
              "REPLICA_SUPERC": ("evals", "dens", "docc", "phisc", "energy", "doubles", "imp"),
              "GENERAL_SUPERC": ("evals", "dens", "docc", "phisc", "energy", "doubles", "imp")}
    for name in checks:
        d = os.path.join(REF, name)
        g = {"source": f"test/src/{name}", "inputs": inputs(os.path.join(d, "inputED.in"))}
        # every *.check of the directory (the tuple above names those the tests used first)
        for f in sorted(os.listdir(d)):
            if f.endswith(".check"):
                g[f[:-6]] = read(os.path.join(d, f))
        for chk in checks[name]:
            assert chk in g, (name, chk)
        # the same goldens are asserted with ED_READ_UMATRIX=T (umatrix.restart) and with the
        # operators added at run time (set_twobody_hk in the driver): keep both operator lists
        um = [f for f in sorted(os.listdir(d)) if f.startswith("umatrix") and f.endswith(".restart")]
        g["umatrix"] = umatrix_lines(os.path.join(d, um[0])) if um else []
        drv = [f for f in os.listdir(d) if f.endswith(".f90")]
        g["twobody_hk"] = twobody_hk_calls(os.path.join(d, drv[0])) if drv else []
        with open(os.path.join(HERE, name.lower() + ".json"), "w") as f:
            json.dump(g, f, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
