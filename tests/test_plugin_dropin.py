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
    from tests import ref_cli
    why = ref_cli.explain_failure(r.returncode)
    if why:
        pytest.fail(why)         # (loud: the prebuilt reference cannot run on this host's ISA)
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
    "white": ('<bsdf type="twosided" id="white"><bsdf type="diffuse"><rgb name="reflectance" value="0.885809, 0.698859, 0.666422"/></bsdf>'
              '<bsdf type="diffuse"><rgb name="reflectance" value="0.2, 0.2, 0.8"/></bsdf></bsdf>'),
}


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_materials(tmp_path):
    """Same, with plastic / conductor / roughconductor / twosided BSDFs instantiated by the reference's own plugins and
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


def _matpreview_xml(integrator, width=171, height=128, spp=16):
    """The SHIPPED resources/data/scenes/matpreview/matpreview.xml (envmap + checkerboard-textured plane + plastic +
    serialized meshes with texture coordinates, gaussian filter); only the integrator type, the film size and the
    sample count are edited."""
    d = os.path.join(ROOT, "oracle", "_ref", "scenes", "matpreview")
    x = open(os.path.join(d, "matpreview.xml")).read()
    x = x.replace('<integrator type="path"/>', f'<integrator type="{integrator}"/>')
    x = x.replace('name="sample_count" value="64"', f'name="sample_count" value="{spp}"')
    x = x.replace('name="width" value="683"', f'name="width" value="{width}"').replace('name="height" value="512"', f'name="height" value="{height}"')
    x = x.replace('value="envmap.exr"', f'value="{d}/envmap.exr"').replace('value="matpreview.serialized"', f'value="{d}/matpreview.serialized"')
    x = x.replace('<string name="pixel_format" value="rgb" />', '<string name="pixel_format" value="rgb" /><string name="component_format" value="float32"/>'
                  '<string name="file_format" value="pfm"/>')
    return x


needs_matpreview = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "scenes", "matpreview", "matpreview.xml")),
                                      reason="oracle/_ref/scenes/matpreview missing (oracle/build_ref.sh)")


@needs_ref
@needs_matpreview
@pytest.mark.gpu
def test_plugin_dropin_shipped_matpreview(tmp_path):
    """BASELINE configs[2]'s scene as shipped: the reference loads it (serialized meshes, EXR envmap, checkerboard,
    plastic), b2path converts and renders it on the GPU; compared with the reference's own `path` render."""
    from tests import golden_cases as G
    r0, ref_out = _run(_matpreview_xml("path"), tmp_path, "ref_mp")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(_matpreview_xml("b2path"), tmp_path, "b2_mp")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    # gaussian filter: atomics order differs from the reference's block merge order -> float rounding only
    assert s["median_rel"] < 5e-6 and s["rel_l2_unflipped"] < 1e-4 and s["flipped_pixels"] <= max(20, s["pixels"] // 100), s


def _write_pfm(path, img):
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write((b"PF" if img.ndim == 3 else b"Pf") + b"\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(img[::-1], "<f4").tobytes())


def _uv_obj(src, dst):
    out = []
    for line in open(src):
        cur = line.strip()
        if cur.startswith("v "):
            x, y, z = [float(t) for t in cur.split()[1:4]]
            out += [line if line.endswith("\n") else line + "\n", "vt %r %r\n" % (x / 500.0 + y / 700.0, z / 500.0 + y / 900.0)]
        elif cur.startswith("f "):
            toks = []
            for t in cur.split()[1:]:
                p = t.split("/")
                toks.append(f"{p[0]}/{p[0]}" + (f"/{p[2]}" if len(p) > 2 and p[2] else ""))
            out.append("f " + " ".join(toks) + "\n")
        else:
            out.append(line)
    open(dst, "w").writelines(out)


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_bitmap_textures(tmp_path):
    """Bitmap textures loaded by the reference (PFM -> BitmapTextureImpl), meshes with texture coordinates loaded by
    its OBJ loader; the adapter recovers filter_type / wrap_mode by probing (they are not exposed by traverse())."""
    import re
    from tests import golden_cases as G
    y, x = np.mgrid[0:6, 0:8].astype(np.float32)
    _write_pfm(str(tmp_path / "rgb.pfm"), np.clip(np.stack([0.15 + 0.1 * x, 0.9 - 0.12 * y, 0.2 + 0.05 * ((x * 3 + y * 5) % 7)], -1), 0, 1).astype(np.float32))
    y, x = np.mgrid[0:7, 0:5].astype(np.float32)
    _write_pfm(str(tmp_path / "gray.pfm"), (0.1 + 0.125 * ((x * 2 + y * 3) % 7)).astype(np.float32))

    def tex(fn, name, filt, wrap, sx, sy):
        return (f'<texture type="bitmap" name="{name}"><string name="filename" value="{tmp_path}/{fn}"/><string name="filter_type" value="{filt}"/>'
                f'<string name="wrap_mode" value="{wrap}"/><boolean name="raw" value="true"/><transform name="to_uv"><scale x="{sx}" y="{sy}"/>'
                f'<translate x="-0.25" y="0.5"/></transform></texture>')
    mats = {"white": f'<bsdf type="diffuse" id="white">{tex("rgb.pfm", "reflectance", "bilinear", "clamp", 1.5, 1.5)}</bsdf>',
            "green": f'<bsdf type="diffuse" id="green">{tex("gray.pfm", "reflectance", "nearest", "repeat", 3.0, 2.0)}</bsdf>',
            "box": f'<bsdf type="plastic" id="box">{tex("rgb.pfm", "diffuse_reflectance", "nearest", "mirror", 2.0, 3.0)}<float name="int_ior" value="1.6"/></bsdf>'}

    def xml(integrator):
        s = _scene_xml(integrator)
        for k, v in mats.items():
            s = re.sub(r'<bsdf type="diffuse" id="%s">.*?</bsdf>' % k, v, s, count=1)
        for fn in set(re.findall(r'value="([^"]+\.obj)"', s)):
            dst = str(tmp_path / ("uv_" + os.path.basename(fn)))
            _uv_obj(fn, dst)
            s = s.replace(fn, dst)
        return s
    r0, ref_out = _run(xml("path"), tmp_path, "ref_t")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(xml("b2path"), tmp_path, "b2_t")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 20, s


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_direct_integrator(tmp_path):
    """`<integrator type="direct">` (shading_samples = 1) vs `<integrator type="b2direct">`: the reference's direct
    integrator consumes the sampler exactly like `path` with max_depth = 2, which is what b2direct runs on the GPU."""
    import re
    from tests import golden_cases as G

    def xml(integrator):
        s = _scene_xml("path").replace('<integrator type="path">', f'<integrator type="{integrator}">')
        return re.sub(r'<integer name="max_depth" value="[^"]*"/>', "", s)
    r0, ref_out = _run(xml("direct"), tmp_path, "ref_d")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(xml("b2direct"), tmp_path, "b2_d")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 20, s


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_direct_integrator_sample_counts(tmp_path):
    """`<integrator type="direct">` with emitter_samples = 3, bsdf_samples = 2 vs `b2direct` (direct.cpp:106-196 proper)."""
    import re
    from tests import golden_cases as G

    def xml(integrator):
        s = _scene_xml("path").replace('<integrator type="path">', f'<integrator type="{integrator}"><integer name="emitter_samples" value="3"/>'
                                       '<integer name="bsdf_samples" value="2"/>')
        return re.sub(r'<integer name="max_depth" value="[^"]*"/>', "", s)
    r0, ref_out = _run(xml("direct"), tmp_path, "ref_d2")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(xml("b2direct"), tmp_path, "b2_d2")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 20, s


@needs_ref
@pytest.mark.gpu
def test_plugin_dropin_stratified_sampler(tmp_path):
    """`<sampler type="stratified">` instantiated by the reference, recognised by the adapter and run on the GPU."""
    from tests import golden_cases as G

    def xml(integrator):
        return _scene_xml(integrator).replace('<sampler type="independent">', '<sampler type="stratified">')
    r0, ref_out = _run(xml("path"), tmp_path, "ref_s")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(xml("b2path"), tmp_path, "b2_s")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 20, s


SHAPES_XML = """
<shape type="shapegroup" id="grp">
  <shape type="obj"><string name="filename" value="{mesh_dir}/cbox_smallbox.obj"/><ref id="green"/></shape>
</shape>
<shape type="instance"><ref id="grp"/><transform name="to_world"><translate x="150" y="0" z="-60"/></transform></shape>
<shape type="instance"><ref id="grp"/><transform name="to_world"><scale value="0.5"/><rotate y="1" angle="30"/><translate x="60" y="300" z="250"/></transform></shape>
<shape type="sphere"><point name="center" x="400" y="90" z="150"/><float name="radius" value="90"/>
  <bsdf type="roughdielectric"><string name="distribution" value="ggx"/><float name="alpha" value="0.15"/><float name="int_ior" value="1.5"/><float name="ext_ior" value="1.0"/></bsdf></shape>
<shape type="rectangle"><transform name="to_world"><scale x="60" y="40"/><rotate x="1" angle="90"/><translate x="278" y="540" z="120"/></transform>
  <ref id="white"/><emitter type="area"><rgb name="radiance" value="6, 7, 9"/></emitter></shape>
<emitter type="directional"><vector name="direction" x="0.3" y="-1" z="0.6"/><rgb name="irradiance" value="1.5, 1.2, 0.9"/></emitter>
"""


@needs_ref
def test_plugin_converts_instances_and_analytic_shapes_cpu(tmp_path):
    """CPU box: the adapter walks Instance / ShapeGroup / Sphere / Rectangle / DirectionalEmitter / RoughDielectric objects of the
    loaded scene and reaches b2_scene_create (which validates the description and then reports the missing device)."""
    from mitsuba2_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    mesh_dir = os.path.join(ROOT, "oracle", "_ref", "scenes", "cbox", "meshes")
    xml = _scene_xml("b2path").replace("</scene>", SHAPES_XML.format(mesh_dir=mesh_dir) + "</scene>")
    r, out = _run(xml, tmp_path, "b2_shapes")
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr), r.stdout + r.stderr


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("device_bvh", [False, True])
def test_plugin_dropin_instances_and_analytic_shapes(tmp_path, device_bvh):
    """BASELINE configs[4]'s ingredients from a Mitsuba scene file: shapegroup + instances (scaled / rotated), an analytic
    sphere with a rough dielectric, an emitting rectangle and a directional emitter — loaded by the reference, converted by
    the adapter (also with the BVHs built on the device), compared with the reference's own render."""
    from tests import golden_cases as G
    mesh_dir = os.path.join(ROOT, "oracle", "_ref", "scenes", "cbox", "meshes")

    def xml(integrator):
        extra = '<boolean name="device_bvh" value="true"/>' if (device_bvh and integrator == "b2path") else ""
        return _scene_xml(integrator, extra=extra).replace("</scene>", SHAPES_XML.format(mesh_dir=mesh_dir) + "</scene>")
    r0, ref_out = _run(xml("path"), tmp_path, "ref_sh")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, b2_out = _run(xml("b2path"), tmp_path, "b2_sh")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(b2_out), _read_pfm(ref_out))
    # (a rough-dielectric sphere: every lobe decision within an ulp of its threshold is a flip; 43 of 4096 pixels measured)
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 100, s


ACCEL = os.path.join(ROOT, "oracle", "_ref", "dist_b2accel")
needs_accel = pytest.mark.skipif(not os.path.exists(os.path.join(ACCEL, "libmitsuba-render.so")),
                                 reason="oracle/_ref/dist_b2accel not built (integration/build_accel_shim.sh)")


def _run_accel(xml, tmp_path, name, threads=2):
    """The reference CLI with the shim build of libmitsuba-render.so first on the library path (Scene::accel_* -> libb2pt)."""
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = ACCEL + ":" + DIST + ":" + env.get("LD_LIBRARY_PATH", "")
    sp = str(tmp_path / f"{name}.xml"); out = str(tmp_path / f"{name}.pfm")
    open(sp, "w").write(xml)
    r = subprocess.run([os.path.join(DIST, "mitsuba"), "-m", "scalar_rgb", "-t", str(threads), "-o", out, sp], env=env, capture_output=True, text=True)
    return r, out


@needs_ref
@needs_accel
def test_accel_shim_fails_loudly_without_gpu(tmp_path):
    from mitsuba2_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    r, out = _run_accel(_scene_xml("path", res=16, spp=1), tmp_path, "acc")
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr), r.stdout + r.stderr


@needs_ref
@needs_accel
@pytest.mark.gpu
def test_accel_shim_keeps_path_on_the_cpu(tmp_path):
    """integration/scene_b2pt.inl: the reference's OWN `path` integrator runs on the CPU; only Scene::ray_intersect /
    ray_test go to the GPU (one b2_stage_intersect / b2_stage_occluded call per ray — the boundary of
    include/mitsuba/render/scene.h:184-205, not a fast path). Same image as the unmodified reference up to closest-hit
    rounding (the well-posed cbox: ceiling patch removed)."""
    from tests import golden_cases as G
    import re
    xml = _scene_xml("path", res=32, spp=4)
    src = re.search(r'value="([^"]+cbox_ceiling\.obj)"', xml).group(1)
    dst = str(tmp_path / "ceiling_nopatch.obj")
    lines = open(src).read().splitlines()
    faces = [ln for ln in lines if ln.startswith("f ")]
    open(dst, "w").write("\n".join([ln for ln in lines if not ln.startswith("f ")] + faces[:1]) + "\n")
    xml = xml.replace(src, dst)
    r0, ref_out = _run(xml, tmp_path, "ref_acc")
    assert r0.returncode == 0, r0.stdout + r0.stderr
    r1, acc_out = _run_accel(xml, tmp_path, "b2_acc")
    assert r1.returncode == 0, r1.stdout + r1.stderr
    s = G.parity_stats(_read_pfm(acc_out), _read_pfm(ref_out))
    assert s["median_rel"] < 2e-6 and s["rel_l2_unflipped"] < 1e-5 and s["flipped_pixels"] <= 10, s
