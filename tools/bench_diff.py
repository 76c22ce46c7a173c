"""Side-by-side view of two bench.py JSON lines (per-kernel ms/step), e.g. pairs on vs off."""
import json, sys

def load(p):
    try:
        return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        return None

a, b = load(sys.argv[1]), load(sys.argv[2]) if len(sys.argv) > 2 else None
if a is None:
    sys.exit("no bench line in " + sys.argv[1])
print("ms/step", a["ms_per_step"], b["ms_per_step"] if b else "-", " hot path", a.get("hot_path_ms_per_step"),
      b.get("hot_path_ms_per_step") if b else "-", " e2e", a["e2e"]["value"] if a.get("e2e") else None)
ka = a.get("hot_path_kernels", {})
kb = b.get("hot_path_kernels", {}) if b else {}
for k, v in ka.items():
    w = kb.get(k)
    print(f"{k:26s} {v['ms_per_step']:.4f} {('%.4f' % w['ms_per_step']) if w else '   -  '}  {v['achieved']:8.1f} {v['unit']:8s} frac {v['frac']:.3f}"
          + (f"  (other {w['frac']:.3f})" if w else ""))
