"""Drop-in test of the reference-side binding (integration/b2path.cpp): the UNMODIFIED reference
CLI loads the scene with its own XML / plugin / OBJ code; only `<integrator type="path">` is
swapped for `<integrator type="b2path">` (plugins/b2path.so -> libb2pt.so -> GPU). Needs
oracle/_ref (built by oracle/build_ref.sh + integration/build_plugin.sh; travels to the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIST = os.path.join(ROOT, "oracle", "_ref", "dist")
sys.path.insert(0, ROOT)


def _scene_xml(integrator, res=64, spp=16, extra=""):
    import bench
    xml = bench._ref_cbox_xml(res, spp, 0, 6)
    return xml.replace('<integrator type="path">', f'<integrator type="{integrator}">{extra}')


def _run(xml, tmp_path, name):
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = DIST + ":" + env.get("LD_LIBRARY_PATH", "")
    sp = str(tmp_path / f"{name}.xml"); out = str(tmp_path / f"{name}.pfm")
    open(sp, "w").write(xml)
    r = subprocess.run([os.path.join(DIST, "mitsuba"), "-m", "scalar_rgb", "-o", out, sp], env=env, capture_output=True, text=True)
    return r, out


def _read_pfm(fn):
    with open(fn, "rb") as f:
        f.readline(); w, h = map(int, f.readline().split()); scale = float(f.readline())
        a = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    return np.ascontiguousarray(a[::-1]).astype(np.float32)


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(DIST, "plugins", "b2path.so")),
                               reason="oracle/_ref/dist/plugins/b2path.so not built (integration/build_plugin.sh)")


@needs_ref
def test_plugin_loads_and_fails_loudly_without_gpu(tmp_path):
    """CPU box: the reference's PluginManager dlopen()s b2path.so and instantiates the class;
    without a CUDA device render() must throw (no CPU fallback)."""
    from mitsuba2_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    r, out = _run(_scene_xml("b2path"), tmp_path, "b2")
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_matches_reference_path(tmp_path):
    from tests import golden_cases as G
    r0, ref_out = _run(_scene_xml("path"), tmp_path, "ref")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(_scene_xml("b2path"), tmp_path, "b2")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    assert "b2path:" in r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    # shipped cbox has coplanar geometry (DESIGN.md 'Float parity'): flip-robust statistics
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 20, s


MATERIALS = {
    "box": '<bsdf type="plastic" id="box"><rgb name="diffuse_reflectance" value="0.1, 0.27, 0.36"/><float name="int_ior" value="1.9"/></bsdf>',
    "green": '<bsdf type="conductor" id="green"><rgb name="eta" value="0.2, 0.92, 1.1"/><rgb name="k" value="3.9, 2.45, 2.14"/></bsdf>',
    "red": '<bsdf type="roughconductor" id="red"><string name="distribution" value="ggx"/><float name="alpha" value="0.15"/></bsdf>',
}


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_materials(tmp_path):
    """Same, with plastic / conductor / roughconductor BSDFs instantiated by the reference's own plugins and
    converted by the adapter (parameters_changed() values taken from the reference's functions)."""
    import re
    from tests import golden_cases as G

    def xml(integrator):
        x = _scene_xml(integrator)
        for k, v in MATERIALS.items():
            x = re.sub(r'<bsdf type="diffuse" id="%s">.*?</bsdf>' % k, v, x, count=1)
        return x
    r0, ref_out = _run(xml("path"), tmp_path, "ref_m")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(xml("b2path"), tmp_path, "b2_m")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 20, s
