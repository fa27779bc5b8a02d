"""On-disk scene INSTANCES.

The reference generates its worlds from an unseeded `thread_rng` inside `Application::new`
(/root/reference/src/application.rs:132-199) and never stores them, so a render cannot be repeated on the same world.
Here a scene instance — the constructor-call tree of :mod:`scene` plus camera and background (application.rs:201-211) —
is written to ONE `.npz` file: a JSON node table (textures, materials, hittables in creation order, children and shared
materials referenced by index) and the bulk arrays (perlin tables, decoded image bytes) as npz members.  Loading
rebuilds the same descriptor objects with the same sharing, so :func:`scene.emit` issues the identical builder-call
sequence: ids, BVH ordering and the flattened op stream are bit-identical (tests/test_scene_io.py).
"""
from __future__ import annotations

import json
from typing import Any, Dict, List

import numpy as np

from . import scene as S

FORMAT = "hrt-scene"
VERSION = 1


def _f3(v):
    return [float(v[0]), float(v[1]), float(v[2])]


class _Writer:
    def __init__(self):
        self.tex: List[dict] = []
        self.mat: List[dict] = []
        self.obj: List[dict] = []
        self.arrays: Dict[str, np.ndarray] = {}
        self._t: Dict[int, int] = {}
        self._m: Dict[int, int] = {}
        self._o: Dict[int, int] = {}

    def _array(self, a: np.ndarray) -> str:
        key = f"a{len(self.arrays)}"
        self.arrays[key] = np.ascontiguousarray(a)
        return key

    def texture(self, t) -> int:
        if id(t) in self._t:
            return self._t[id(t)]
        if isinstance(t, S.SolidColor):
            n = {"k": "solid", "color": _f3(t.color)}
        elif isinstance(t, S.CheckerTexture):
            odd = self.texture(t.odd)
            even = self.texture(t.even)
            n = {"k": "checker", "odd": odd, "even": even}
        elif isinstance(t, S.NoiseTexture):
            p = t.noise
            n = {"k": "noise", "scale": float(t.scale), "ranvec": self._array(np.asarray(p.random_vectors, dtype=np.float32)),
                 "px": self._array(np.asarray(p.permutation_x, dtype=np.uint32)),
                 "py": self._array(np.asarray(p.permutation_y, dtype=np.uint32)),
                 "pz": self._array(np.asarray(p.permutation_z, dtype=np.uint32))}
        elif isinstance(t, S.ImageTexture):
            n = {"k": "image", "data": None if t.data is None else self._array(np.asarray(t.data, dtype=np.uint8))}
        else:
            raise TypeError(f"not a texture: {t!r}")
        self.tex.append(n)
        self._t[id(t)] = len(self.tex) - 1
        return self._t[id(t)]

    def material(self, m) -> int:
        if id(m) in self._m:
            return self._m[id(m)]
        if isinstance(m, S.Lambertian):
            n = {"k": "lambertian", "albedo": self.texture(m.albedo)}
        elif isinstance(m, S.Metal):
            n = {"k": "metal", "albedo": _f3(m.albedo), "fuzz": float(m.fuzz)}
        elif isinstance(m, S.Dielectric):
            n = {"k": "dielectric", "ior": float(m.index_of_refraction)}
        elif isinstance(m, S.DiffuseLight):
            n = {"k": "diffuse_light", "emit": self.texture(m.emit)}
        else:
            raise TypeError(f"not a material: {m!r}")
        self.mat.append(n)
        self._m[id(m)] = len(self.mat) - 1
        return self._m[id(m)]

    def hittable(self, h) -> int:
        if id(h) in self._o:
            return self._o[id(h)]
        if isinstance(h, S.Sphere):
            n = {"k": "sphere", "center": _f3(h.center), "radius": float(h.radius), "mat": self.material(h.material)}
        elif isinstance(h, S.MovingSphere):
            n = {"k": "moving_sphere", "c0": _f3(h.center_start), "c1": _f3(h.center_end), "t0": float(h.time_start),
                 "t1": float(h.time_end), "radius": float(h.radius), "mat": self.material(h.material)}
        elif isinstance(h, S.Rect):
            n = {"k": "rect", "plane": int(h.plane), "a0": float(h.a0), "a1": float(h.a1), "b0": float(h.b0), "b1": float(h.b1),
                 "kk": float(h.k), "mat": self.material(h.material)}
        elif isinstance(h, S.Cuboid):
            n = {"k": "cuboid", "min": _f3(h.box_min), "max": _f3(h.box_max), "mat": self.material(h.material)}
        elif isinstance(h, S.Translation):
            n = {"k": "translate", "child": self.hittable(h.hittable), "d": _f3(h.displacement)}
        elif isinstance(h, S.Rotation):
            n = {"k": "rotate", "axis": int(h.axis), "child": self.hittable(h.hittable), "angle": float(h.angle)}
        elif isinstance(h, S.ConstantMedium):
            boundary = self.hittable(h.boundary)
            n = {"k": "medium", "boundary": boundary, "density": float(h.density), "tex": self.texture(h.texture)}
        elif isinstance(h, S.List):
            n = {"k": "list", "objects": [self.hittable(o) for o in h.objects]}
        elif isinstance(h, S.BvhNode):
            n = {"k": "bvh", "objects": [self.hittable(o) for o in h.objects], "t0": float(h.time_start), "t1": float(h.time_end)}
        else:
            raise TypeError(f"not a hittable: {h!r}")
        self.obj.append(n)
        self._o[id(h)] = len(self.obj) - 1
        return self._o[id(h)]


def save_scene(spec: S.SceneSpec, path: str) -> None:
    """Write one scene instance (world + camera + background) to `path` (.npz)."""
    w = _Writer()
    root = w.hittable(spec.world)
    c = spec.camera
    doc = {"format": FORMAT, "version": VERSION, "name": spec.name, "root": root, "background": _f3(spec.background),
           "camera": {"look_from": _f3(c.look_from), "look_at": _f3(c.look_at), "fov": float(c.fov), "aperture": float(c.aperture),
                      "focus_dist": float(c.focus_dist), "time_0": float(c.time_0), "time_1": float(c.time_1)},
           "textures": w.tex, "materials": w.mat, "hittables": w.obj}
    blob = np.frombuffer(json.dumps(doc).encode("utf-8"), dtype=np.uint8)
    with open(path, "wb") as f:  # np.savez would append ".npz" to a bare name
        np.savez_compressed(f, scene_json=blob, **w.arrays)


def load_scene(path: str) -> S.SceneSpec:
    """Read a scene instance written by :func:`save_scene`; raises ValueError on a foreign or newer file."""
    with np.load(path, allow_pickle=False) as z:
        if "scene_json" not in z.files:
            raise ValueError(f"{path}: not a {FORMAT} file")
        doc = json.loads(bytes(z["scene_json"]).decode("utf-8"))
        arrays = {k: z[k] for k in z.files if k != "scene_json"}
    if doc.get("format") != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} file")
    if int(doc.get("version", -1)) != VERSION:
        raise ValueError(f"{path}: {FORMAT} version {doc.get('version')} is not supported (this build reads version {VERSION})")

    tex: List[Any] = []
    for n in doc["textures"]:  # creation order: children always precede their users
        k = n["k"]
        if k == "solid":
            tex.append(S.SolidColor(tuple(n["color"])))
        elif k == "checker":
            tex.append(S.CheckerTexture(tex[n["odd"]], tex[n["even"]]))
        elif k == "noise":
            tex.append(S.NoiseTexture(n["scale"], S.PerlinNoise(arrays[n["ranvec"]], arrays[n["px"]], arrays[n["py"]], arrays[n["pz"]])))
        elif k == "image":
            tex.append(S.ImageTexture(None if n["data"] is None else arrays[n["data"]]))
        else:
            raise ValueError(f"{path}: unknown texture kind {k!r}")
    mat: List[Any] = []
    for n in doc["materials"]:
        k = n["k"]
        if k == "lambertian":
            mat.append(S.Lambertian(tex[n["albedo"]]))
        elif k == "metal":
            mat.append(S.Metal(tuple(n["albedo"]), n["fuzz"]))
        elif k == "dielectric":
            mat.append(S.Dielectric(n["ior"]))
        elif k == "diffuse_light":
            mat.append(S.DiffuseLight(tex[n["emit"]]))
        else:
            raise ValueError(f"{path}: unknown material kind {k!r}")
    obj: List[Any] = []
    for n in doc["hittables"]:
        k = n["k"]
        if k == "sphere":
            obj.append(S.Sphere(tuple(n["center"]), n["radius"], mat[n["mat"]]))
        elif k == "moving_sphere":
            obj.append(S.MovingSphere(tuple(n["c0"]), tuple(n["c1"]), n["t0"], n["t1"], n["radius"], mat[n["mat"]]))
        elif k == "rect":
            obj.append(S.Rect(n["plane"], n["a0"], n["a1"], n["b0"], n["b1"], n["kk"], mat[n["mat"]]))
        elif k == "cuboid":
            obj.append(S.Cuboid(tuple(n["min"]), tuple(n["max"]), mat[n["mat"]]))
        elif k == "translate":
            obj.append(S.Translation(obj[n["child"]], tuple(n["d"])))
        elif k == "rotate":
            obj.append(S.Rotation(n["axis"], obj[n["child"]], n["angle"]))
        elif k == "medium":
            obj.append(S.ConstantMedium(obj[n["boundary"]], n["density"], tex[n["tex"]]))
        elif k == "list":
            obj.append(S.List([obj[i] for i in n["objects"]]))
        elif k == "bvh":
            obj.append(S.BvhNode([obj[i] for i in n["objects"]], n["t0"], n["t1"]))
        else:
            raise ValueError(f"{path}: unknown hittable kind {k!r}")
    c = doc["camera"]
    cam = S.Camera(tuple(c["look_from"]), tuple(c["look_at"]), c["fov"], c["aperture"], c["focus_dist"], c["time_0"], c["time_1"])
    return S.SceneSpec(doc["name"], obj[doc["root"]], cam, tuple(doc["background"]))
