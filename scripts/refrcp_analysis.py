#!/usr/bin/env python
"""How much of the oracle-vs-reference noise floor is the reciprocal? (CPU only, AVX-512 host; VERDICT round 1, "attack C1".)

The reference evaluates rcp / rsqrt as vrcp14ss / vrsqrt14ss + one Newton step (accurate to ~1 ulp, not correctly rounded);
the oracle and the CUDA kernels use the correctly rounded 1/x and 1/sqrt(x); the reference calls glibc's sinf / cosf / expf / logf /
atan2f, the oracle and the kernels share polynomial twins. `oracle/Makefile` builds ANALYSIS variants of the oracle that execute the
reference's reciprocal sequences (libb2oracle_refrcp.so), call glibc (libb2oracle_reflibm.so), or both. This script renders every
golden case with each of them and prints flipped pixels / rel-L2 / median relative error against the reference's image."""
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

LIBS = (("as_shipped", "libb2oracle.so"), ("reference_rcp", "libb2oracle_refrcp.so"), ("glibc_libm", "libb2oracle_reflibm.so"),
        ("rcp_and_libm", "libb2oracle_refboth.so"))
subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")] + [l for _, l in LIBS], check=True, capture_output=True)
res = {}
for tag, lib in LIBS:
    env = dict(os.environ, B2O_LIB=os.path.join(ROOT, "oracle", lib))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker"], env=env, capture_output=True, text=True, check=True)
    res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
tags = [t for t, _ in LIBS]
print(f"{'case':44s} " + "  ".join(f"{t:>28s}" for t in tags) + "     (flipped pixels | rel-L2 | median rel. error)")
for name in res[tags[0]]:
    print(f"{name:44s} " + "  ".join(f"{res[t][name]['flipped_pixels']:6d} {res[t][name]['rel_l2']:.2e} {res[t][name]['median_rel']:.1e}" .rjust(28) for t in tags))
for t in tags:
    v = res[t]
    print(f"{t:16s} total flipped pixels {sum(c['flipped_pixels'] for c in v.values()):5d}   cases with median error 0: {sum(1 for c in v.values() if c['median_rel'] == 0.0):2d} / {len(v)}"
          f"   mean of the medians {sum(c['median_rel'] for c in v.values()) / len(v):.2e}")
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
