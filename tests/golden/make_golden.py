"""Extracts the reference's only recorded results -- the cell outputs of /root/reference/doc/examples.ipynb -- into
tests/golden/notebook_outputs.json.  Run in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference has no test fixtures (test/runtests.jl:4-6 is empty) and cannot be executed here (no Julia), so these
statistical results plus the analytic identities of SURVEY.md section 4 are what pins the oracle.
"""
import json
import os
import re

SRC = "/root/reference/doc/examples.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_outputs.json")


def cell_text(cell):
    out = []
    for o in cell.get("outputs", []):
        out.append("".join(o.get("text", [])) if "text" in o else "".join(o.get("data", {}).get("text/plain", [])))
    return "".join(out)


def main():
    nb = json.load(open(SRC))
    cells = nb["cells"]
    gold = {"source": "lgresista/SemiClassicalMC.jl doc/examples.ipynb (recorded cell outputs, Julia 1.8.2)", "model": {}}
    src4 = "".join(cells[4]["source"])
    m = re.search(r"Isv, Is, Iv = ([-\d.]+), ([-\d.]+), ([-\d.]+)", src4)
    gold["model"] = {"Isv": float(m.group(1)), "Is": float(m.group(2)), "Iv": float(m.group(3)), "lattice": "triangular", "n": 1,
                     "J_diag_order": "[0, Iv, Iv, Iv, Is, Isv, Isv, Isv, Is, Isv, Isv, Isv, Is, Isv, Isv, Isv]"}
    src11 = "".join(cells[11]["source"])
    run = {k: float(re.search(rf"\b{k}\s*=\s*([\d.e+-]+)", src11).group(1)) for k in ("T", "T_i", "N_therm", "N_measure", "measurement_rate")}
    run["L"] = 6
    gold["metropolis"] = {"params": run, "thermalization_log": [], "measurement_acceptance": []}
    for idx in (11, 15, 17, 19):
        txt = cell_text(cells[idx])
        for sw, T, R, sg in re.findall(r"(\d+)/\d+ thermalization sweeps\. T = ([\d.e-]+), R = ([\d.e-]+), σ = ([\d.e-]*\d)", txt):
            gold["metropolis"]["thermalization_log"].append({"cell": idx, "sweep": int(sw), "T": float(T), "R": float(R), "sigma": float(sg)})
        gold["metropolis"]["measurement_acceptance"] += [float(x) for x in re.findall(r"measurement sweeps at ([\d.]+) acceptance rate", txt)]
    txt13 = cell_text(cells[13])
    means = {}
    for name, val in re.findall(r'f\["means/([^"]+)"\]\)\[\] = ([-\d.e]+)', txt13):
        means[name] = float(val)
    gold["metropolis"]["means"] = means
    txt25 = cell_text(cells[25])
    src25 = "".join(cells[25]["source"])
    ann = {}
    for k in ("L", "T_i", "T_fac", "N_max", "N_o", "N_per_T", "min_acc_per_site", "min_accrate", "σ_min"):
        mm = re.search(rf"\b{k}\s*=\s*([\d.e+-]+)", src25)
        if mm:
            ann[k] = float(mm.group(1))
    gold["annealing"] = {"params": ann,
                         "checkpoints": [{"sweep": int(s), "T": float(T), "E": float(E), "R": float(R), "sigma": float(sg)} for s, T, E, R, sg in
                                         re.findall(r"checkpoint after (\d+)/\d+ sweeps, T = ([\d.e-]+), E = ([-\d.e]+), R = ([\d.e-]+), σ = ([\d.e-]*\d)", txt25)],
                         "stop_sweep": int(re.search(r"Minimal acceptance rate reached at (\d+)/", txt25).group(1)),
                         "optimization_E": [float(x) for x in re.findall(r"optimization sweeps done\. E = ([-\d.e]+)", txt25)],
                         "exact_E_per_site": -3.6}
    json.dump(gold, open(OUT, "w"), indent=1, ensure_ascii=False)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
