"""Summarise gpurun_out/parity_errors.json (written by a `pytest -m gpu` session, tests/_tol.py): the largest measured
error per family of checks and dtype, as a markdown table.   usage: python tools/parity_summary.py [IN.json] [OUT.md]"""
import collections, json, re, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_errors.json"
rows = json.load(open(src))


def family(what):
    w = re.sub(r"\b(n|N|M|T|dim|batch|scale|eps|matrix|rank|rigorous)[ =]*[-0-9.e]+", "", what)
    w = re.sub(r"\s+", " ", w).strip(" ,")
    return w


agg = collections.OrderedDict()
for r in rows:
    k = (family(r["what"]), r["dtype"], r["named_shape"])
    a = agg.setdefault(k, {"n": 0, "max": 0.0, "bar": r["bar"], "worst": ""})
    a["n"] += 1
    if r["err"] >= a["max"]:
        a["max"], a["worst"] = r["err"], r["what"]
out = ["| check | dtype | cases | largest error | bar | worst case |", "|---|---|---:|---:|---:|---|"]
for (fam, dt, named), a in sorted(agg.items(), key=lambda kv: (kv[0][1], -kv[1]["max"] / kv[1]["bar"])):
    out.append(f"| {fam}{'' if named else ' (stress bound)'} | {dt} | {a['n']} | {a['max']:.2e} | {a['bar']:.0e} | {a['worst']} |")
text = f"{len(rows)} recorded checks; worst error / bar = {max(r['err'] / r['bar'] for r in rows):.3f}\n\n" + "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
else:
    print(text)
