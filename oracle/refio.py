"""TEST INFRASTRUCTURE — talks to oracle/_ref/<f32|f64>/ref_driver (the unmodified reference build).

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may import this.

Scene file (little endian), read by oracle/ref_driver.cpp:
    char[8]  "SHKZIN02"
    int32    nx, ny, nz, solid_mode, fluid_mode, repeat, pad, pad
    float64  dx, dt, surface_tension, band, current_volume, target_volume
    float32  u[(nx+1)*ny*nz], v[nx*(ny+1)*nz], w[nx*ny*(nz+1)]          (x fastest)
    uint8    face-active masks, same three shapes
    float32  solid_raw[(nx+1)(ny+1)(nz+1)]     iff solid_mode == 1
    float32  fluid_raw[nx*ny*nz]               iff fluid_mode == 1
Result file:
    char[8]  "SHKZOUT2"; int32 nx,ny,nz,sizeof(Real); float64 ms_last, ms_mean
    6 + (6 if fractions) dense blocks: int32 w,h,d,has_active; float64 values[w*h*d]; uint8 active[w*h*d] (if has_active)
    order: u, v, w, pressure, fluid, solid, <int32 has_fractions>, areas[3], rhos[3]
"""
from __future__ import annotations

import os
import re
import struct
import subprocess
import tempfile
from dataclasses import dataclass
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_dir(real: str = "f32") -> str:
    return os.path.join(HERE, "_ref", real)


def ref_available(real: str = "f32") -> bool:
    return os.path.isfile(os.path.join(ref_dir(real), "ref_driver"))


def write_scene(path: str, scene, repeat: int = 1, current_volume: float = 0.0, target_volume: float = 0.0):
    assert scene.zrange == (0, scene.nz), "the reference runs whole grids only"
    fluid_mode = 0 if scene.fluid_raw is None else 1
    with open(path, "wb") as f:
        f.write(b"SHKZIN02")
        f.write(struct.pack("<8i", scene.nx, scene.ny, scene.nz, scene.solid_mode, fluid_mode, repeat, 0, 0))
        f.write(struct.pack("<6d", scene.dx, scene.dt, scene.surface_tension, scene.band, current_volume, target_volume))
        for v in scene.vel:
            f.write(np.ascontiguousarray(v, dtype=np.float32).tobytes())
        for a in scene.vel_active:
            f.write(np.ascontiguousarray(a, dtype=np.uint8).tobytes())
        if scene.solid_mode == 1:
            f.write(np.ascontiguousarray(scene.solid_raw, dtype=np.float32).tobytes())
        if fluid_mode == 1:
            f.write(np.ascontiguousarray(scene.fluid_raw, dtype=np.float32).tobytes())


@dataclass
class RefResult:
    vel: list
    vel_active: list
    pressure: np.ndarray
    pressure_active: np.ndarray
    fluid: np.ndarray
    solid: np.ndarray
    areas: Optional[list]
    rhos: Optional[list]
    iterations: int
    reresid: float
    ms_project: float
    sizeof_real: int
    phase_ms: dict
    stdout: str
    records: Optional[dict] = None   # records=True: console::write record name -> list of values (one per call)


def _read_block(f):
    w, h, d, has_active = struct.unpack("<4i", f.read(16))
    n = w * h * d
    values = np.frombuffer(f.read(8 * n), dtype=np.float64).reshape(d, h, w)
    active = np.frombuffer(f.read(n), dtype=np.uint8).reshape(d, h, w) if has_active else None
    return values, active


def read_result(path: str):
    with open(path, "rb") as f:
        assert f.read(8) == b"SHKZOUT2"
        nx, ny, nz, sizeof_real = struct.unpack("<4i", f.read(16))
        ms_last, ms_mean = struct.unpack("<2d", f.read(16))
        blocks = [_read_block(f) for _ in range(6)]
        (has_fractions,) = struct.unpack("<i", f.read(4))
        areas = rhos = None
        if has_fractions:
            areas = [_read_block(f)[0] for _ in range(3)]
            rhos = [_read_block(f)[0] for _ in range(3)]
    return dict(vel=[b[0] for b in blocks[:3]], vel_active=[b[1] for b in blocks[:3]],
                pressure=blocks[3][0], pressure_active=blocks[3][1], fluid=blocks[4][0], solid=blocks[5][0],
                areas=areas, rhos=rhos, ms_project=ms_mean, sizeof_real=sizeof_real)


_TIME = r"([0-9.]+) (msec|sec|minutes|hours|days)"
_UNIT_MS = {"msec": 1.0, "sec": 1e3, "minutes": 6e4, "hours": 3.6e6, "days": 8.64e7}
_ANSI = re.compile(r"\x1b\[[0-9;]*m")


def _phase_times(text: str) -> dict:
    """The reference prints its scoped_timer values as console text (macpressuresolver3.cpp:82,201,237,269,271)."""
    out = {}
    pats = {
        "solid_fluid_fractions": r"Precomputing solid and fluid fractions\.\.\.Done\. Took " + _TIME,
        "build_highres_linsystem": r"Building the high-res linear system \[Lhs\] and \[rhs\]\.\.\.Done\. Took " + _TIME,
        "linsolve": r"Reresid=[-+0-9.eE]+\. Took " + _TIME,
        "update_velocity": r"Updating the velocity\.\.\.Done\. Took " + _TIME,
        "projection": r"Projection done\. Took " + _TIME,
    }
    for key, pat in pats.items():
        m = None
        for m in re.finditer(pat, text):
            pass
        if m:
            out[key] = float(m.group(1)) * _UNIT_MS[m.group(2)]
    return out


def run_reference(scene, real: str = "f32", flags: Optional[dict] = None, dump_fractions: bool = False,
                  repeat: int = 1, threads: Optional[int] = None, projection: Optional[str] = None,
                  current_volume: float = 0.0, target_volume: float = 0.0, extra_lib_dirs=(),
                  timeout: Optional[float] = None, records: bool = False, extrapolate: Optional[int] = None, skip_project: bool = False,
                  advect: Optional[str] = None, advection: Optional[str] = None) -> RefResult:
    """One project() call of the reference (or of any drop-in module named by `projection`)."""
    d = ref_dir(real)
    if not ref_available(real):
        raise RuntimeError(f"oracle/_ref/{real}/ref_driver not built (run `make -C oracle ref` where /root/reference exists)")
    with tempfile.TemporaryDirectory(prefix="shkzref_") as tmp:
        fin, fout = os.path.join(tmp, "scene.bin"), os.path.join(tmp, "result.bin")
        write_scene(fin, scene, repeat=repeat, current_volume=current_volume, target_volume=target_volume)
        argv = [os.path.join(d, "ref_driver"), f"in={fin}", f"out={fout}"]
        if dump_fractions:
            argv.append("DumpFractions=1")
        if projection:
            argv.append(f"Projection={projection}")
        if threads:
            argv.append(f"Threads={threads}")
        if records:
            argv.append(f"RecordDir={tmp}")
        if extrapolate is not None:   # the reference's macutility3::extrapolate_and_constrain_velocity after (or, skip_project, instead of) the projection
            argv.append(f"RefExtrapolate={int(extrapolate)}")
        if skip_project:
            argv.append("RefSkipProject=1")
        if advect:   # INSTEAD of the projection: one advect_vector ("vector") / advect_scalar ("density", "levelset") call of the module `advection` names
            argv.append(f"RefAdvect={advect}")   # (default: the reference's macadvection3); the advected scalar comes back in the pressure slot
        if advection:
            argv.append(f"Advection={advection}")
        for k, v in (flags or {}).items():
            argv.append(f"{k}={v}")
        env = dict(os.environ)
        env["LD_LIBRARY_PATH"] = os.pathsep.join([d, *extra_lib_dirs, env.get("LD_LIBRARY_PATH", "")])
        proc = subprocess.run(argv, env=env, cwd=tmp, capture_output=True, text=True, timeout=timeout)
        text = _ANSI.sub("", proc.stdout)
        if proc.returncode != 0 or not os.path.isfile(fout):
            raise RuntimeError(f"ref_driver failed ({proc.returncode}):\n{text[-4000:]}\n{proc.stderr[-4000:]}")
        res = read_result(fout)
        rec = None
        if records:
            rec = {}
            rdir = os.path.join(tmp, "record")
            for fn in sorted(os.listdir(rdir)) if os.path.isdir(rdir) else []:
                if fn.endswith(".out"):
                    with open(os.path.join(rdir, fn)) as fh:
                        rec[fn[:-4]] = [float(line.split()[1]) for line in fh if len(line.split()) >= 2]
    m = None
    for m in re.finditer(r"Took (\d+) iterations, Reresid=([-+0-9.]+(?:[eE][-+]?\d+)?|nan|inf|-nan)", text):
        pass
    iterations = int(m.group(1)) if m else -1
    reresid = float(m.group(2)) if m else float("nan")
    return RefResult(iterations=iterations, reresid=reresid, phase_ms=_phase_times(text), stdout=text, records=rec, **res)
