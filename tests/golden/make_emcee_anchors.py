#!/usr/bin/env python
"""Extract the REAL-emcee posterior numbers the reference holds in its stored notebook outputs
(run where /root/reference exists):

    python tests/golden/make_emcee_anchors.py

emcee is a third-party, unpinned dependency of the reference (requirements.txt:2) whose source is
neither vendored nor installable here; the only outputs of the real library the reference tree
holds are the cell outputs of docs/tutorials/*.ipynb and docs/tutorials/quickstart_results.csv.
This script parses them (no number is typed by hand) into tests/golden/emcee_anchors.json,
together with the exact configuration of the cell that produced each of them, so that the
sampler restatement (oracle/) and the CUDA sampler can be z-tested against emcee's own numbers.

Every anchor records the notebook cell and its line range in the .ipynb file.
"""
import json
import os
import re

REF = "/root/reference/docs/tutorials"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emcee_anchors.json")


def load_nb(name):
    path = os.path.join(REF, name)
    with open(path) as f:
        text = f.read()
    return json.loads(text), text.splitlines()


def line_of(lines, needle, start=0):
    for i in range(start, len(lines)):
        if needle in lines[i]:
            return i + 1
    raise KeyError(needle)


def cell_text(cell):
    out = []
    for o in cell.get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain")
        if t:
            out.append("".join(t))
    return "\n".join(out)


def cell_latex(cell):
    out = []
    for o in cell.get("outputs", []):
        t = o.get("data", {}).get("text/latex")
        if t:
            out.append("".join(t))
    return out


def find_cell(nb, needle):
    for i, c in enumerate(nb["cells"]):
        if c["cell_type"] == "code" and needle in "".join(c["source"]):
            return i, c
    raise KeyError(needle)


def latex_mean_std(cell):
    mean, std, dec = [], [], []
    for s in cell_latex(cell):
        m = re.search(r":\s*(-?[0-9.]+)\s*\\pm\s*([0-9.]+)", s)
        mean.append(float(m.group(1)))
        std.append(float(m.group(2)))
        dec.append(len(m.group(1).split(".")[1]))
    return mean, std, dec


def main():
    anchors = []

    # ---- quickstart: PeltonColeCole n_modes=1, K389172, headers=9, 32 x 2000, discard 500 thin 2 (unseeded) ----
    nb, lines = load_nb("quickstart.ipynb")
    _, c_fit = find_cell(nb, "headers=9")
    src = "".join(c_fit["source"])
    assert "nwalkers=32" in src and "nsteps=2000" in src and "SIP-K389172" in src
    _, c_chain = find_cell(nb, "get_chain(discard=500, thin=2, flat=True)")
    assert "(24000, 4)" in cell_text(c_chain)
    _, c_ms = find_cell(nb, "get_param_std(chain=chain)")
    names, mean, std = [], [], []
    for ln in cell_text(c_ms).strip().splitlines():
        m = re.match(r"(\w+): (-?[0-9.]+) \+/- ([0-9.]+)", ln)
        names.append(m.group(1)); mean.append(float(m.group(2))); std.append(float(m.group(3)))
    with open(os.path.join(REF, "quickstart_results.csv")) as f:
        hdr = f.readline().strip().split(",")
        pct = [[float(v) for v in ln.split(",")] for ln in f.read().strip().splitlines()]
    assert hdr == names
    # the csv is the full-precision copy of the notebook's printed `results`
    _, c_res = find_cell(nb, "print(results)")
    printed = [float(v) for v in re.findall(r"-?\d+\.\d+", cell_text(c_res))]
    assert all(abs(a - b) < 5e-9 for a, b in zip(printed, [v for r in pct for v in r]))
    anchors.append(dict(
        name="quickstart_cc1", model="colecole", file="SIP-K389172", headers=9, n_modes=1, nwalkers=32, nsteps=2000,
        discard=500, thin=2, bounds_edits={}, param_names=names,
        source=f"docs/tutorials/quickstart.ipynb:{line_of(lines, 'r0: 1.02413')}-{line_of(lines, 'c1: 0.50193')}; "
               "docs/tutorials/quickstart_results.csv:2-4",
        mean=mean, std=std, mean_decimals=[5] * 4, std_decimals=[5] * 4, p=[2.5, 50, 97.5], pct=pct))

    # ---- decomposition: PD Debye poly_deg 4, 32 x 1000, discard 500, six files (np.random.seed(42)) ----
    nb, lines = load_nb("decomposition.ipynb")
    _, c_par = find_cell(nb, "results_debye = {}")
    src = "".join(c_par["source"])
    assert "'nwalkers': 32" in src and "'nsteps': 1000" in src and "'c_exp': 1" in src and "'poly_deg': 4" in src
    _, c_tab = find_cell(nb, "m_debye = get_m(df_debye")
    txt = cell_text(c_tab)
    rows, tot = {}, {}
    for ln in txt.splitlines():
        m = re.match(r"(SIP-K\d+)\s+(.*?)\s*\\?$", ln.strip())
        if not m:
            continue
        vals = [float(v) for v in m.group(2).split()]
        if len(vals) == 6:
            rows[m.group(1)] = vals
        elif len(vals) == 1:
            tot[m.group(1)] = vals[0]
    assert len(rows) == 6 and len(tot) == 6
    l0, l1 = line_of(lines, "SIP-K389170  0.989771"), line_of(lines, "SIP-K389176  0.545386")
    for f in sorted(rows):
        anchors.append(dict(
            name=f"decomp_debye_p4_{f[-7:]}", model="decomp", file=f, headers=1, poly_deg=4, c_exp=1.0, nwalkers=32,
            nsteps=1000, discard=500, thin=1, bounds_edits={}, param_names=["r0", "a0", "a1", "a2", "a3", "a4"],
            source=f"docs/tutorials/decomposition.ipynb:{l0}-{l1} (get_param_mean(discard=500), total_m)",
            mean=rows[f], mean_decimals=[6] * 6, total_m=tot[f], total_m_decimals=6))

    # ---- pelton: CC2, K389174, 64 x 1000, edited bounds, discard 500 thin 10 (seed 42) ----
    nb, lines = load_nb("pelton.ipynb")
    _, c_b = find_cell(nb, "model.params.update(log_tau1=[-5, 5], log_tau2=[-15, -10])")
    _, c_ms = find_cell(nb, "values = model.get_param_mean(chain)")
    _, c_ch = find_cell(nb, "chain = model.get_chain(discard=500, thin=10, flat=True)")
    mean, std, dec = latex_mean_std(c_ms)
    assert len(mean) == 7
    anchors.append(dict(
        name="pelton_cc2_K389174", model="colecole", file="SIP-K389174", headers=1, n_modes=2, nwalkers=64, nsteps=1000,
        discard=500, thin=10, bounds_edits={"log_tau1": [-5, 5], "log_tau2": [-15, -10]},
        param_names=["r0", "m1", "m2", "log_tau1", "log_tau2", "c1", "c2"],
        source=f"docs/tutorials/pelton.ipynb:{line_of(lines, 'rho_0: 1.010')}-{line_of(lines, 'c_2: 0.608')} "
               f"(bounds edited at :{line_of(lines, 'log_tau2=[-15, -10]')})",
        mean=mean, std=std, mean_decimals=dec, std_decimals=dec))

    # ---- dias: K389172, 32 x 1000, edited bounds, discard 500 (seed 42) ----
    nb, lines = load_nb("dias.ipynb")
    _, c_b = find_cell(nb, "model.params.update(eta=[0, 25], log_tau=[-15, -5])")
    _, c_ms = find_cell(nb, "values = model.get_param_mean(discard=500)")
    mean, std, dec = latex_mean_std(c_ms)
    assert len(mean) == 5
    anchors.append(dict(
        name="dias_K389172", model="dias", file="SIP-K389172", headers=1, nwalkers=32, nsteps=1000, discard=500, thin=1,
        bounds_edits={"eta": [0, 25], "log_tau": [-15, -5]}, param_names=["r0", "m", "log_tau", "eta", "delta"],
        source=f"docs/tutorials/dias.ipynb:{line_of(lines, 'rho_0: 1.023')}-{line_of(lines, 'delta: 0.707')} "
               f"(bounds edited at :{line_of(lines, 'eta=[0, 25]')})",
        mean=mean, std=std, mean_decimals=dec, std_decimals=dec))

    with open(OUT, "w") as f:
        json.dump({"note": "real-emcee outputs stored in the reference's notebooks; generated by make_emcee_anchors.py",
                   "anchors": anchors}, f, indent=1)
    for a in anchors:
        print(a["name"], a["source"])


if __name__ == "__main__":
    main()
