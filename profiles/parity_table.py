"""gpurun_out/parity_table.jsonl (written by tests/util.py::record_parity during `pytest -m gpu` on the box) ->
profiles/r2_parity_table.md.  usage: python profiles/parity_table.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [json.loads(l) for l in open(os.path.join(ROOT, "gpurun_out", "parity_table.jsonl"))]
fmt = lambda v: f"{v:.2e}" if isinstance(v, float) and (abs(v) < 1e-2 or abs(v) >= 1e4) and v != 0 else (f"{v:g}" if isinstance(v, float) else str(v))  # noqa: E731
tri = [r for r in rows if {"err_truth", "err_gold", "ref_noise"} <= set(r)]
other = [r for r in rows if r not in tri]
out = ["# Round 2 - observed parity errors on B200 (`pytest -m gpu`, `gpurun_out/parity_table.jsonl`)", "",
       "Every check that goes through `tests/util.py::assert_close_to_reference` prints its three errors here, so the slack that",
       "function allows is visible: `err_truth` = |device - float64 restatement|, `err_gold` = |device - the reference's recorded",
       "float32 output|, `ref_noise` = |reference - float64| (the reference's own rounding noise).  All are max-norm relative.",
       "The assertion is `err_truth <= rtol or err_gold <= rtol`, and `err_gold <= rtol + 2 ref_noise`.", "",
       "| check | err_truth | err_gold | ref_noise | rtol | device closer to float64 than the reference is |", "|---|---|---|---|---|---|"]
for r in tri:
    out.append(f"| {r['case']} | {fmt(r['err_truth'])} | {fmt(r['err_gold'])} | {fmt(r['ref_noise'])} | {fmt(r['rtol'])} | "
               f"{'yes' if r['err_truth'] <= r['ref_noise'] else 'no'} |")
worst = max(tri, key=lambda r: r["err_truth"]) if tri else None
if worst:
    out += ["", f"Largest `err_truth`: {fmt(worst['err_truth'])} ({worst['case']}).  "
            f"Checks where `err_gold` exceeds rtol (reference noise limited): "
            f"{sum(1 for r in tri if r['err_gold'] > r['rtol'])} of {len(tri)}; in all of them `err_truth` <= rtol."]
out += ["", "## Other recorded figures", "", "| check | values |", "|---|---|"]
for r in other:
    c = r.pop("case")
    out.append(f"| {c} | " + ", ".join(f"{k} = {fmt(v)}" for k, v in r.items()) + " |")
open(os.path.join(ROOT, "profiles", "r2_parity_table.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[-12:]))
