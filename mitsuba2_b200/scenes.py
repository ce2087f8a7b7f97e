"""Scene builders for the benchmark / parity configurations (BASELINE.json configs).

Each builder returns abi.SceneBuffers — the flat description the C-ABI consumes. `host` is
any object exposing the b2_host_* helpers (look_at, parse_fov, perspective_sensor, film_box,
film_gaussian): mitsuba2_b200.api.Host for the product, tests/oracle_lib.OracleHost for the
oracle, so both sides can be fed bit-identical inputs or each derive their own.
"""
import os

import numpy as np

from . import abi

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# resources/data/scenes/cbox/fragments/bsdfs-rgb.xml, cbox-rgb.xml
CBOX_BSDFS = {
    "box": (0.45, 0.30, 0.90), "white": (0.885809, 0.698859, 0.666422),
    "red": (0.570068, 0.0430135, 0.0443706), "green": (0.105421, 0.37798, 0.076425),
    "light": (0.936461, 0.740433, 0.705267),
}
CBOX_RADIANCE = (18.387, 10.9873, 2.75357)


def cbox(host, width=256, height=256, rfilter="box", materials=None, meshes_npz=None, lift=0.0, instances=None, drop_meshes=(), envmap=None):
    """cbox-rgb.xml (resources/data/scenes/cbox): 8 OBJ meshes / 38 triangles, one area
    emitter, diffuse rgb BSDFs, perspective sensor fov 39.3077 ('smaller' axis).

    materials: optional {bsdf_id: callable(scene_buffers) -> bsdf index} overriding a BSDF.
    Mesh data comes from tests/golden/cbox_meshes.npz (generated from the reference's OBJ
    files by tests/golden/make_golden.py).
    """
    sb = abi.SceneBuffers()
    z = np.load(meshes_npz or os.path.join(_GOLDEN, "cbox_meshes.npz"))
    bsdf_ids = {}
    for name, refl in CBOX_BSDFS.items():
        if materials and name in materials:
            bsdf_ids[name] = materials[name](sb)
        else:
            bsdf_ids[name] = sb.add_diffuse(refl)
    i = 0
    while f"m{i}_positions" in z:
        b = str(z[f"m{i}_bsdf"])
        pos = z[f"m{i}_positions"]
        if lift and b == "box":     # <translate y=lift/> on the two boxes (applied in fp32 like obj.cpp:172)
            pos = (pos + np.array([0, lift, 0], np.float32)).astype(np.float32)
        if i in drop_meshes:
            i += 1
            continue
        if instances is not None and b == "box" and i == 7:
            i += 1
            continue                      # instanced variant: the large box is dropped ...
        mi = sb.add_mesh(pos, z[f"m{i}_faces"], z[f"m{i}_normals"], bsdf_ids[b])
        if b == "light":
            sb.add_area_emitter(mi, CBOX_RADIANCE)
        if instances is not None and b == "box" and i == 6:
            g = sb.add_shapegroup([mi])   # ... and the small box becomes a shapegroup instantiated several times
            for m in instances:
                sb.add_instance(g, m)
        i += 1
    if envmap is not None:        # dict(rgb=[H,W,3], scale=, to_world=4x4): appended after the shapes (XML order)
        sb.add_envmap(envmap["rgb"], envmap.get("scale", 1.0), envmap.get("to_world"))
    sb.film = host.film_box(width, height) if rfilter == "box" else host.film_gaussian(width, height, 0.5)
    to_world = host.look_at((278, 273, -800), (278, 273, -799), (0, 1, 0))
    fov_x = host.parse_fov(39.3077, "smaller", width, height)
    sb.sensor = host.perspective_sensor(to_world, fov_x, 10.0, 2800.0, sb.film)
    return sb
