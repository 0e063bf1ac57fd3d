"""Index-level emulation of k_gl_ws (thread mapping, padded smem layout, geometry) against the einsum."""
import numpy as np
def geom(G, Ig, Hg, CPT):
    need = (Hg + CPT - 1) // CPT
    tpg = 1
    while tpg < need: tpg *= 2
    gcta = min(256 // tpg, G)
    assert G % gcta == 0
    rs = 256 // (gcta * tpg)
    row_bytes = gcta * (Ig + 4) * 4
    R = min(40 * 1024 // row_bytes, 32)
    R = max(R, rs); R -= R % rs
    return tpg, gcta, rs, R, 2 * R * row_bytes
def run(M, G, Ig, Hg, CPT, seed=0):
    rng = np.random.default_rng(seed)
    I, H = G * Ig, G * Hg
    x = rng.standard_normal((M, I)).astype(np.float32)
    W = rng.standard_normal((G, Ig, Hg)).astype(np.float32)
    ref = np.einsum("mgk,gkn->mgn", x.reshape(M, G, Ig), W).reshape(M, H)
    tpg, gcta, rs, R, smem = geom(G, Ig, Hg, CPT)
    assert smem <= 100 * 1024 and gcta * tpg * rs == 256
    GS, XP = Ig + 4, gcta * (Ig + 4)
    y = np.full((M, H), np.nan, np.float32)
    ntiles = (M + R - 1) // R
    gy = G // gcta
    for by in range(gy):
        for tile in range(ntiles):
            m0 = tile * R
            buf = np.full(R * XP, np.nan, np.float32)
            cpr = gcta * (Ig // 4)
            for idx in range(R * cpr):
                r, c = divmod(idx, cpr); cg, k4 = divmod(c, Ig // 4)
                dst = r * XP + cg * GS + k4 * 4
                m = m0 + r
                buf[dst:dst + 4] = x[m, by * gcta * Ig + cg * Ig + k4 * 4: by * gcta * Ig + cg * Ig + k4 * 4 + 4] if m < M else 0
            for tid in range(256):
                per_slot = gcta * tpg
                slot, within = divmod(tid, per_slot); gl, tc = divmod(within, tpg)
                g = by * gcta + gl; col0 = tc * CPT
                for r in range(slot, R, rs):
                    m = m0 + r
                    if m >= M: break
                    xs = buf[r * XP + gl * GS: r * XP + gl * GS + Ig]
                    for c in range(CPT):
                        if col0 + c < Hg:
                            assert np.isnan(y[m, g * Hg + col0 + c]), "written twice"
                            y[m, g * Hg + col0 + c] = np.dot(xs.astype(np.float64), W[g, :, col0 + c].astype(np.float64))
    assert not np.isnan(y).any(), "unwritten outputs"
    err = np.abs(y - ref).max()
    print(f"G={G} Ig={Ig} Hg={Hg} CPT={CPT}: tpg={tpg} gcta={gcta} rs={rs} R={R} smem={smem} grid.y={gy} max err {err:.2e}")
    assert err < 1e-4
for (G, Ig, Hg, CPT) in [(32, 96, 16, 1), (16, 32, 16, 2), (8, 64, 32, 1), (16, 16, 32, 2), (16, 16, 60, 4)]:
    run(37, G, Ig, Hg, CPT)
