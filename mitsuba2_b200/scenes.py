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


def cbox(host, width=256, height=256, rfilter="box", materials=None, meshes_npz=None, lift=0.0, instances=None, drop_meshes=(), envmap=None,
         texcoords=None, crop=None):
    """cbox-rgb.xml (resources/data/scenes/cbox): 8 OBJ meshes / 38 triangles, one area
    emitter, diffuse rgb BSDFs, perspective sensor fov 39.3077 ('smaller' axis).

    materials: optional {bsdf_id: callable(scene_buffers) -> bsdf index} overriding a BSDF.
    texcoords: optional {mesh index: [n_vertices, 2] uv array}.
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
    sb.bsdf_ids = bsdf_ids          # cbox BSDF id -> index of its (front) record
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
        mi = sb.add_mesh(pos, z[f"m{i}_faces"], z[f"m{i}_normals"], bsdf_ids[b], texcoords=texcoords.get(i) if texcoords else None)
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
    if crop:                      # (offset_x, offset_y, width, height), film.cpp:17-27
        sb.film.crop_offset[:] = crop[:2]; sb.film.crop_size[:] = crop[2:]
    to_world = host.look_at((278, 273, -800), (278, 273, -799), (0, 1, 0))
    fov_x = host.parse_fov(39.3077, "smaller", width, height)
    sb.sensor = host.perspective_sensor(to_world, fov_x, 10.0, 2800.0, sb.film)
    return sb


# ---------------------------------------------------------------------------------------------
# "veach-mis"-class scene (BASELINE configs[3]): the meshes of the reference's doc scene
# (resources/data/docs/scenes/include/veach_mis.xml) are not shipped, so the scene is synthesised:
# four tilted plates (rough conductors of graded roughness + one dielectric), a diffuse floor and
# four spherical area emitters (tessellated icospheres) of graded size / radiance.
# ---------------------------------------------------------------------------------------------
def _icosphere(center, radius, level=1):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = list(v / np.linalg.norm(v, axis=1, keepdims=True))
    for _ in range(level):
        cache = {}; nf = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (v[a] + v[b]) * 0.5
                v.append(m / np.linalg.norm(m)); cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    n = np.array(v, np.float64)
    pos = (np.array(center, np.float64) + radius * n).astype(np.float32)
    return pos, n.astype(np.float32), np.array(f, np.uint32)


def _quad(p0, p1, p2, p3):
    pos = np.array([p0, p1, p2, p3], np.float32)
    e0, e1 = pos[1] - pos[0], pos[3] - pos[0]
    n = np.cross(e0.astype(np.float64), e1.astype(np.float64)); n = (n / np.linalg.norm(n)).astype(np.float32)
    return pos, np.tile(n, (4, 1)).astype(np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


VEACH_EMITTERS = [  # centre, radius, radiance — graded like veach_mis.xml:20-70
    ((-3.75, 0.0, 0.0), 0.0333 * 3, (901.803, 901.803, 901.803)),
    ((-1.25, 0.0, 0.0), 0.1 * 3, (100.0, 100.0, 100.0)),
    ((1.25, 0.0, 0.0), 0.3 * 3, (11.1111, 11.1111, 11.1111)),
    ((3.75, 0.0, 0.0), 0.9 * 1.2, (1.23457, 1.23457, 1.23457)),
]


def veach_mis_meshes():
    """-> list of dicts(name, positions, normals, faces, material) — deterministic."""
    out = []
    out.append(dict(name="floor", material=("diffuse", (0.4, 0.4, 0.4)),
                    mesh=_quad((-10, -4.15, -10), (-10, -4.15, 10), (10, -4.15, 10), (10, -4.15, -10))))
    out.append(dict(name="backwall", material=("diffuse", (0.4, 0.4, 0.4)),
                    mesh=_quad((-10, -4.15, -2), (10, -4.15, -2), (10, 10, -2), (-10, 10, -2))))
    plates = [(-2.6, 5.0, 0.35, ("roughconductor", "ggx", 0.01)), (-3.0, 3.9, 0.55, ("roughconductor", "beckmann", 0.05)),
              (-3.4, 2.8, 0.8, ("dielectric", 1.5)), (-3.8, 1.6, 1.1, ("roughconductor", "ggx", 0.1))]
    for i, (y, z, tilt, mat) in enumerate(plates):
        dy, dz = 0.55 * np.sin(tilt), 0.55 * np.cos(tilt)
        out.append(dict(name=f"plate{i}", material=mat,
                        mesh=_quad((-4, y - dy, z + dz), (4, y - dy, z + dz), (4, y + dy, z - dz), (-4, y + dy, z - dz))))
    for i, (c, r, rad) in enumerate(VEACH_EMITTERS):
        out.append(dict(name=f"emitter{i}", material=("diffuse", (0.0, 0.0, 0.0)), mesh=_icosphere(c, r, 1), radiance=rad))
    return out


def veach_mis(host, width=1280, height=720, rfilter="box"):
    sb = abi.SceneBuffers()
    for m in veach_mis_meshes():
        mat = m["material"]
        if mat[0] == "diffuse":
            b = sb.add_diffuse(mat[1])
        elif mat[0] == "dielectric":
            b = sb.add_dielectric(mat[1], 1.0)
        else:
            b = sb.add_roughconductor(alpha=mat[2], distribution=mat[1])
        pos, nrm, faces = m["mesh"]
        mi = sb.add_mesh(pos, faces, nrm, b)
        if "radiance" in m:
            sb.add_area_emitter(mi, m["radiance"])
    sb.film = host.film_box(width, height) if rfilter == "box" else host.film_gaussian(width, height, 0.5)
    to_world = host.look_at((0, 2, 15), (0, -2, 2.5), (0, 1, 0))
    fov_x = host.parse_fov(28.0, "x", width, height)
    sb.sensor = host.perspective_sensor(to_world, fov_x, 0.1, 100.0, sb.film)
    return sb


# ---------------------------------------------------------------------------------------------
# BASELINE configs[4]: procedural instanced mesh — one displaced-grid mesh in a shapegroup,
# instantiated on an n^3 lattice with a per-instance rotation drawn from a seeded generator;
# one 2-triangle area emitter above; diffuse 0.5.
# ---------------------------------------------------------------------------------------------
def instanced_grid(host, width=2048, height=2048, grid=71, lattice=10, seed=1234, rfilter="box", flatten=False):
    """grid x grid vertices -> 2 (grid-1)^2 triangles per instance (71 -> 9800), lattice^3 instances.
    flatten: instead of instancing, ONE mesh holding every instance's transformed copy (lattice^3 * 9800 triangles:
    10 M for the defaults) — the 'flattened C5' the BVH builders are timed on."""
    sb = abi.SceneBuffers()
    diffuse = sb.add_diffuse((0.5, 0.5, 0.5))
    light_b = sb.add_diffuse((0.0, 0.0, 0.0))
    # displaced unit grid in the xy plane, z = 0.05 sin(13 x) cos(17 y)
    u = np.linspace(-0.5, 0.5, grid, dtype=np.float64)
    X, Y = np.meshgrid(u, u, indexing="xy")
    Z = 0.05 * np.sin(13 * X) * np.cos(17 * Y)
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    idx = np.arange(grid * grid).reshape(grid, grid)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel(), idx[1:, :-1].ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.uint32)
    if not flatten:
        gm = sb.add_mesh(pos, faces, None, diffuse)
        g = sb.add_shapegroup([gm])
    flat_pos = []
    rng = np.random.default_rng(seed)
    span = float(lattice) * 1.5
    for i in range(lattice ** 3):
        ix, iy, iz = i % lattice, (i // lattice) % lattice, i // (lattice * lattice)
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4); M[:3, :3] = R
        M[:3, 3] = [(ix + 0.5) * 1.5 - span / 2, (iy + 0.5) * 1.5 - span / 2, (iz + 0.5) * 1.5 - span / 2]
        if flatten:
            M32 = M.astype(np.float32)
            flat_pos.append((pos @ M32[:3, :3].T + M32[:3, 3]).astype(np.float32))
        else:
            sb.add_instance(g, M.astype(np.float32))
    if flatten:
        nv = len(pos); ni = len(flat_pos)
        all_faces = (faces[None, :, :] + (np.arange(ni, dtype=np.uint32) * np.uint32(nv))[:, None, None]).reshape(-1, 3).astype(np.uint32)
        sb.add_mesh(np.concatenate(flat_pos), all_faces, None, diffuse)
    # area emitter above the lattice, facing down
    s = span * 0.35; yl = span / 2 + 1.0
    lp = np.array([[-s, yl, -s], [s, yl, -s], [s, yl, s], [-s, yl, s]], np.float32)
    lm = sb.add_mesh(lp, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), None, light_b)
    sb.add_area_emitter(lm, (40.0, 40.0, 40.0))
    sb.film = host.film_box(width, height) if rfilter == "box" else host.film_gaussian(width, height, 0.5)
    cam = span * 1.4
    to_world = host.look_at((cam * 0.8, cam * 0.55, cam), (0, 0, 0), (0, 1, 0))
    sb.sensor = host.perspective_sensor(to_world, host.parse_fov(40.0, "x", width, height), 0.05, 1000.0, sb.film)
    return sb


# ---------------------------------------------------------------------------------------------
# BASELINE configs[2] class ("matpreview roughconductor GGX + envmap"): the shipped matpreview meshes
# are serialized files only the reference can load (they are exercised through the plugin,
# tests/test_plugin_dropin.py), so the bench workload is synthesised with the same ingredients: a
# tessellated ball (rough conductor, GGX) + an inner diffuse base on a checkerboard-textured plane
# with texture coordinates, lit by a lat-long environment map only.
# ---------------------------------------------------------------------------------------------
def _env_image(w=256, h=128):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w, 3), np.float32)
    img[..., 0] = 0.3 + 0.5 * (y / (h - 1))
    img[..., 1] = 0.4 + 0.3 * np.cos(x / (w - 1) * np.float32(6.2831853)) ** 2
    img[..., 2] = 0.9 - 0.6 * (y / (h - 1))
    img[h // 6:h // 3, w // 5:w // 5 + w // 10] += np.array([30.0, 24.0, 12.0], np.float32)
    img[h // 3:h // 2, (3 * w) // 5:(3 * w) // 5 + w // 8] += np.array([2.0, 5.0, 9.0], np.float32)
    return img.astype(np.float32)


def matball(host, width=1920, height=1080, rfilter="gaussian", level=5, alpha=0.1):
    sb = abi.SceneBuffers()
    tex = sb.add_checkerboard((0.4, 0.4, 0.4), (0.2, 0.2, 0.2), [[8, 0, 0], [0, 8, 0], [0, 0, 1]])
    plane_b = sb.add_diffuse(texture=tex)
    ball_b = sb.add_roughconductor(alpha=alpha, distribution="ggx", eta=(0.2, 0.92, 1.1), k=(3.9, 2.45, 2.14))
    base_b = sb.add_diffuse((0.18, 0.18, 0.18))
    pos, nrm, faces = _quad((-6, 0, -6), (-6, 0, 6), (6, 0, 6), (6, 0, -6))
    sb.add_mesh(pos, faces, nrm, plane_b, texcoords=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    pos, nrm, faces = _icosphere((0.0, 1.0, 0.0), 1.0, level)          # 20 * 4^level triangles (level 5: 20480)
    sb.add_mesh(pos, faces, nrm, ball_b)
    pos, nrm, faces = _icosphere((1.9, 0.45, 0.8), 0.45, max(level - 2, 1))
    sb.add_mesh(pos, faces, nrm, base_b)
    sb.add_envmap(_env_image(), 3.0, np.array([[0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]], np.float32))
    sb.film = host.film_box(width, height) if rfilter == "box" else host.film_gaussian(width, height, 0.5)
    to_world = host.look_at((3.7, 2.6, 3.5), (0.2, 0.8, 0.0), (0, 1, 0))
    sb.sensor = host.perspective_sensor(to_world, host.parse_fov(28.8415, "smaller", width, height), 0.1, 100.0, sb.film)
    return sb
