#!/bin/bash
# compute-sanitizer on small cases: memcheck (all kernels incl. the CSR solver) and racecheck (shared-memory hazards of the TMA sweep:
# stage ring, half-updated-plane ring, in-stage coarse correction)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp
from shiokaze_b200 import MacPressureSolver3, B200CG, scenes
mode = sys.argv[1]
cases = [(scenes.dambreak(40, True), "mixed"), (scenes.smoke_plume(32), "fp32"), (scenes.random_blobs(20, 14, 18, seed=3), "fp64"), (scenes.flip_splash(48), "mixed")]
if mode == "race":
    cases = cases[:2]
for sc, prec in cases:
    S = MacPressureSolver3((sc.nx, sc.ny, sc.nz), sc.dx, Precision=prec, Residual=1e-5)
    out = S.project_scene(sc, surface_tension=0.01)
    print(prec, sc.name, out["result"].iterations, out["result"].converged, out["result"].reresid, flush=True)
    S.close()
if mode == "mem":
    # the sparse host-copy kernels (k_flag_wet_slices / k_pull_slices / k_push_faces, k_store_pressure into host memory): page-locked buffers, two scenes in a row
    import ctypes as C
    from shiokaze_b200 import capi
    L = capi.lib(); held = []
    def pinned(a):
        p = C.c_void_p(); capi.check(L.shkz_b200_host_alloc(a.nbytes, C.byref(p))); held.append(p)
        v = np.frombuffer((C.c_ubyte * a.nbytes).from_address(p.value), dtype=a.dtype).reshape(a.shape); v[...] = a; return v
    S = MacPressureSolver3((72, 72, 72), 1.0 / 72)
    pres, pact = pinned(np.zeros((72, 72, 72), np.float32)), pinned(np.zeros((72, 72, 72), np.uint8))
    for sc in (scenes.dambreak(72, True), scenes.flip_splash(72)):
        _, _, res = S.project(sc.dt, [pinned(v) for v in sc.vel], [pinned(a) for a in sc.vel_active], pinned(sc.solid), pinned(sc.fluid), sc.fluid_levelset, pressure_out=pres, pressure_active_out=pact)
        print("sparse host copies", sc.name, res.iterations, res.converged, res.stats["host_copies"], res.stats["h2d_bytes"], res.stats["d2h_bytes"], flush=True)
    S.close()
    # the advection kernels (csrc/advect.cu): every flag combination on faces and cells, whole-array and sparse host copies (page-locked buffers)
    from shiokaze_b200 import MacAdvection3
    sc = scenes.dambreak(40, True)
    rng = np.random.default_rng(1)
    vel = [np.where(a != 0, rng.standard_normal(v.shape) * 0.1, 0).astype(np.float32) for v, a in zip(sc.vel, sc.vel_active)]
    for flags in ({}, {"MacCormack": "No"}, {"WENO": "Yes"}):
        A = MacAdvection3(sc.shape, sc.dx, **flags)
        out = A.advect_vector(vel, sc.vel_active, sc.fluid, 0.3)
        q = A.advect_scalar(sc.fluid, (np.abs(sc.fluid) < sc.band).astype(np.uint8), vel, sc.vel_active, sc.fluid, 0.3, background=sc.band)
        st = A.advect_vector_inplace([pinned(v) for v in vel], [pinned(a.astype(np.uint8)) for a in sc.vel_active], sc.fluid, 0.3)
        print("advect", flags, float(np.abs(out[1]).max()), float(q.min()), st["host_copies"], st["h2d_bytes"], flush=True)
        A.close()
    for p in held: L.shkz_b200_host_free(p)
    n = 24; I = sp.identity(n); T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    A = (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()
    C = B200CG(Residual=1e-8); x, r = C.solve(A, None, None, np.ones(A.shape[0])); print("csr ell", r.count, r.converged)
    B = sp.random(400, 400, density=0.2, random_state=1, format="csr"); W = (B @ B.T + sp.identity(400)).tocsr()
    x, r = C.solve(W, None, None, np.ones(400)); print("csr wide", r.count, r.converged, r.stats["ell_width"]); C.close()
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_case.py mem > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|csr|mixed|fp32|fp64|sparse host|advect|Error|rror:" gpurun_out/sanitizer_memcheck.log | tail -14
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python /tmp/san_case.py race > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|mixed|fp32|rror:" gpurun_out/sanitizer_racecheck.log | sort | uniq -c | sort -rn | head -12
