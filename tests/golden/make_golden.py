"""Copies the reference's golden *.check values for the fixtures this repo pins into small
JSON files (run HERE, where /root/reference exists; the GPU box only sees the JSON).

    python tests/golden/make_golden.py
"""
import json
import os

REF = "/root/reference/test/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(path):
    with open(path) as f:
        return [float(x) for x in f.read().split()]


def inputs(path):
    out = {}
    with open(path) as f:
        for line in f:
            if "=" not in line:
                continue
            k, v = line.split("!")[0].split("=", 1)
            out[k.strip()] = v.strip()
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
              "HYBRID_SUPERC": ("evals", "dens", "docc", "phisc", "energy", "doubles", "imp")}
    for name in checks:
        d = os.path.join(REF, name)
        g = {"source": f"test/src/{name}", "inputs": inputs(os.path.join(d, "inputED.in"))}
        for chk in checks[name]:
            g[chk] = read(os.path.join(d, chk + ".check"))
        with open(os.path.join(HERE, name.lower() + ".json"), "w") as f:
            json.dump(g, f, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
