"""Scratch: write harvested equilibrium systems (tools/harvest_systems.py) as the binary input of
tools/ubench/nnls_warp.cu: int32 count; per system int32 m, n; f64 b[m] (normalised); f64 R[n][m] (the cone edge
rays, column after column); f64 expected residual r* (oracle/nnls.py, cross-checked against scipy BVLS).
Test infrastructure only.

python tools/dump_systems.py /tmp/systems.pkl /tmp/systems.bin [max_systems]"""
import os, struct, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import solver_lab as SL
from oracle import nnls
from oracle import stability as st

if __name__ == "__main__":
    systems = SL.load(sys.argv[1])
    limit = int(sys.argv[3]) if len(sys.argv) > 3 else len(systems)
    rng = np.random.default_rng(0)
    picked = [systems[i] for i in rng.permutation(len(systems))[:limit]]
    with open(sys.argv[2], "wb") as fh:
        fh.write(struct.pack("<i", len(picked)))
        for A, b, mu, ok, nbk, tag in picked:
            R = st.ray_matrix(A, mu)
            bs = b / np.linalg.norm(b)
            r, _ = nnls.equilibrium_residual_nnls(A, b, mu)
            r_bvls = st.equilibrium_residual(A, b, mu)
            assert abs(r - r_bvls) <= 1e-9 + 1e-8 * r_bvls
            m, n = R.shape
            fh.write(struct.pack("<ii", m, n))
            fh.write(np.ascontiguousarray(bs, dtype="<f8").tobytes())
            fh.write(np.ascontiguousarray(R.T, dtype="<f8").tobytes())
            fh.write(struct.pack("<d", r))
    print(f"{len(picked)} systems -> {sys.argv[2]}")
