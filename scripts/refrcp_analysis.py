#!/usr/bin/env python
"""How much of the oracle-vs-reference noise floor is the reciprocal? (CPU only, AVX-512 host; VERDICT round 1, "attack C1".)

The reference evaluates rcp / rsqrt as vrcp14ss / vrsqrt14ss + one Newton step (accurate to ~1 ulp, not correctly rounded);
the oracle and the CUDA kernels use the correctly rounded 1/x and 1/sqrt(x). `make -C oracle libb2oracle_refrcp.so` builds
an ANALYSIS variant of the oracle that executes the reference's instruction sequences instead. This script renders every golden
case with both oracles and prints, per case, rel-L2 / un-flipped rel-L2 / flipped pixels against the reference's image."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    from tests import golden_cases as G, oracle_lib as O
    out = {}
    for name in G.names():
        sb, rp, ref = G.load(name, O.OracleHost())
        img, st = O.render(sb, rp)
        s = G.parity_stats(O.develop(img), ref)
        out[name] = {k: (float(s[k]) if k != "flipped_pixels" and k != "pixels" else int(s[k])) for k in ("rel_l2", "rel_l2_unflipped", "median_rel", "flipped_pixels", "pixels")}
    print(json.dumps(out))
    sys.exit(0)

subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "libb2oracle.so", "libb2oracle_refrcp.so"], check=True, capture_output=True)
res = {}
for tag, lib in (("correctly_rounded", "libb2oracle.so"), ("reference_rcp", "libb2oracle_refrcp.so")):
    env = dict(os.environ, B2O_LIB=os.path.join(ROOT, "oracle", lib))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, capture_output=True, text=True, check=True)
    res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
rows = []
for name in res["correctly_rounded"]:
    a, b = res["correctly_rounded"][name], res["reference_rcp"][name]
    rows.append((name, a, b))
    print(f"{name:44s} flips {a['flipped_pixels']:5d} -> {b['flipped_pixels']:5d} / {a['pixels']:7d}   rel_l2 {a['rel_l2']:.2e} -> {b['rel_l2']:.2e}   "
          f"unflipped {a['rel_l2_unflipped']:.2e} -> {b['rel_l2_unflipped']:.2e}   median {a['median_rel']:.1e} -> {b['median_rel']:.1e}")
tot_a = sum(a["flipped_pixels"] for _, a, _ in rows); tot_b = sum(b["flipped_pixels"] for _, _, b in rows)
print(f"total flipped pixels over {len(rows)} cases: {tot_a} -> {tot_b}")
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
