"""one-line digest of a bench.py JSON line (used by the profiles/run_gpu_*.sh scripts)"""
import json
import sys

for p in sys.argv[1:]:
    try:
        d = json.load(open(p))
    except Exception as e:  # noqa
        print(p, "unreadable:", e)
        try:
            print(open(p.replace(".json", ".err")).read()[-1500:])
        except Exception:
            pass
        continue
    r = d.get("roofline") or {}
    print(p.split("/")[-1], "N=%s" % d.get("n_gpus"), "%.2f Gp/s" % (d["value"] / 1e9), "%.3f ms" % d["ms_per_step"],
          {k: round(v, 3) for k, v in (r.get("stages_ms") or {}).items() if v > 0.004},
          "frac %.3f pipe %.3f" % (r.get("frac", 0), (r.get("pipeline") or {}).get("frac", 0)),
          "parity", (d.get("parity_vs_oracle") or {}).get("ok"), "run", d.get("run"),
          "e2e %.2f Gp/s" % (d["e2e"]["value"] / 1e9) if d.get("e2e") else "")
    for k, b in (d.get("configs") or {}).items():
        rr = b.get("roofline") or {}
        print("   ", k, "%.2f Gp/s" % (b["value"] / 1e9), ("%.3f ms" % b["ms_per_step"]) if "ms_per_step" in b else "",
              {kk: round(v, 3) for kk, v in (rr.get("stages_ms") or {}).items() if v > 0.004},
              "pipe %.3f" % (rr.get("pipeline") or {}).get("frac", 0) if rr else "",
              "parity", (b.get("parity_vs_oracle") or {}).get("ok"),
              "e2e %.2f Gp/s" % (b["e2e"]["value"] / 1e9) if b.get("e2e") else "")
