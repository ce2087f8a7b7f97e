"""Run the unmodified reference CLI (oracle/_ref/dist/mitsuba, built by oracle/build_ref.sh) — test infrastructure only."""
import os
import re
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIST = os.path.join(ROOT, "oracle", "_ref", "dist")


def read_pfm(fn):
    with open(fn, "rb") as f:
        f.readline(); w, h = map(int, f.readline().split()); scale = float(f.readline())
        a = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    return np.ascontiguousarray(a[::-1]).astype(np.float32)


def available():
    return os.path.exists(os.path.join(DIST, "mitsuba"))


def isa_missing():
    """ISA extensions of the host oracle/_ref was built on (-march=native, recorded by oracle/build_ref.sh) that THIS host
    lacks. Empty list = the prebuilt reference can run here."""
    try:
        need = set(open(os.path.join(ROOT, "oracle", "_ref", "ISA_FLAGS.txt")).read().split())
        have = set()
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                have = set(ln.split(":", 1)[1].split()); break
        return sorted(need - have)
    except Exception:
        return []


def explain_failure(returncode):
    """A loud explanation when the prebuilt CLI dies with SIGILL on a host with another ISA."""
    if returncode in (-4, 132):
        return ("oracle/_ref was built with -march=native on another host and hit an illegal instruction here; this host lacks: "
                + (" ".join(isa_missing()) or "(none of the recorded extensions — unexpected)") + ". Rebuild with oracle/build_ref.sh.")
    return None


def render_xml(xml, threads=None):
    """-> (rgb image, seconds logged by 'Rendering finished. (took ...)', integrator.cpp:174-176) or (None, None)"""
    if not available():
        return None, None
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = DIST + ":" + env.get("LD_LIBRARY_PATH", "")
    with tempfile.TemporaryDirectory() as td:
        sp = os.path.join(td, "s.xml"); out = os.path.join(td, "o.pfm")
        open(sp, "w").write(xml)
        cmd = [os.path.join(DIST, "mitsuba"), "-m", "scalar_rgb"] + (["-t", str(threads)] if threads else []) + ["-o", out, sp]
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3600)
        except Exception:
            return None, None
        if r.returncode != 0 or not os.path.exists(out):
            why = explain_failure(r.returncode)
            if why:
                raise RuntimeError(why)
            return None, None
        m = re.search(r"Rendering finished\. \(took ([0-9.]+)(ms|s|m)", r.stdout + r.stderr)
        t = float(m.group(1)) * {"ms": 1e-3, "s": 1.0, "m": 60.0}[m.group(2)] if m else None
        return read_pfm(out), t


def render_cbox(res, spp, spp_pass, max_depth, threads=None):
    import sys
    sys.path.insert(0, ROOT)
    import bench
    return render_xml(bench._ref_cbox_xml(res, spp, spp_pass, max_depth), threads)
