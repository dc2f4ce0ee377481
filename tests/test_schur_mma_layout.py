"""The operand mapping of the tensor-core Schur pair walk (csrc/pxr_ba_kernels.cuh, schur_pairs_accumulate_fused), restated
in numpy: mma.sync.m8n8k4.f64 takes from lane L = 4*gid + tig the elements A[gid][tig] and B[tig][gid] and leaves
D[gid][2*tig + i] (PTX ISA, "Matrix Fragments for mma.m8n8k4 with .f64"; CuTe's SM80_8x4 / SM80_8x8_Row say the same).
The kernel's index arithmetic — record offset 3*gid + tig, the symmetric-inverse index table, H^-1 in the EVEN columns of the
first B so that its accumulator is the next A fragment, g_p in column 0 — must reproduce T_x W_y^T and T_x g_p.  This test
checks that arithmetic on the CPU; the kernel itself is checked against the oracle by the -m gpu parity tests."""
import numpy as np


def mma_884(a_frag, b_frag, c_frag):
    """one warp-wide m8n8k4: fragments indexed by lane -> D fragments [32][2]"""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    for lane in range(32):
        gid, tig = lane >> 2, lane & 3
        A[gid, tig] = a_frag[lane]
        B[tig, gid] = b_frag[lane]
        C[gid, 2 * tig] = c_frag[lane][0]; C[gid, 2 * tig + 1] = c_frag[lane][1]
    D = A @ B + C
    return [[D[lane >> 2, 2 * (lane & 3)], D[lane >> 2, 2 * (lane & 3) + 1]] for lane in range(32)]


def test_fused_pair_walk_operand_mapping():
    rng = np.random.default_rng(0)
    dcm = 10                                                   # record = dcm rows of 3 doubles; rows >= 8 never read here
    Wx, Wy = rng.normal(size=(dcm, 3)), rng.normal(size=(dcm, 3))
    M = rng.normal(size=(3, 3)); Hinv = M @ M.T + np.eye(3)    # symmetric, stored as (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    hsym = np.array([Hinv[0, 0], Hinv[0, 1], Hinv[0, 2], Hinv[1, 1], Hinv[1, 2], Hinv[2, 2]])
    gp = rng.normal(size=3)
    a, b, h, g = np.zeros(32), np.zeros(32), np.zeros(32), np.zeros(32)
    for lane in range(32):
        gid, tig = lane >> 2, lane & 3
        ld = tig < 3
        e = gid * 3 + tig if ld else 0                         # the 24 lanes with tig < 3 cover the record's 24 doubles
        a[lane] = Wx.reshape(-1)[e] if ld else 0.0
        b[lane] = Wy.reshape(-1)[e] if ld else 0.0
        ldh = ld and gid < 6 and gid % 2 == 0                  # B[k][2j] = Hinv[k][j]
        lo, hi = min(gid >> 1, tig), max(gid >> 1, tig)
        hidx = (hi if lo == 0 else (2 + hi if lo == 1 else 5)) if ldh else 0
        h[lane] = hsym[hidx] if ldh else 0.0
        g[lane] = gp[tig] if (ld and gid == 0) else 0.0        # B2[k][0] = g_p[k]
    assert sorted(set(3 * (l >> 2) + (l & 3) for l in range(32) if (l & 3) < 3)) == list(range(24))
    zero = [[0.0, 0.0]] * 32
    t = [d[0] for d in mma_884(a, h, zero)]                    # accumulator element 0 = D[gid][2*tig] = T[gid][tig]
    T = Wx[:8] @ Hinv
    for lane in range(32):
        gid, tig = lane >> 2, lane & 3
        assert abs(t[lane] - (T[gid, tig] if tig < 3 else 0.0)) < 1e-13
    D = mma_884(t, b, zero)
    R = mma_884(t, g, zero)
    want = T @ Wy[:8].T
    for lane in range(32):
        gid, tig = lane >> 2, lane & 3
        assert abs(D[lane][0] - want[gid, 2 * tig]) < 1e-12 and abs(D[lane][1] - want[gid, 2 * tig + 1]) < 1e-12
        if tig == 0:
            assert abs(R[lane][0] - T[gid] @ gp) < 1e-12


def test_rows_and_columns_beyond_an_images_block_stay_where_they_are():
    """an image with fewer than 8 columns leaves garbage rows in its records: an MMA never mixes rows of A or columns of B,
    so the entries the sink writes (row < dcx, column < dcy) do not see it"""
    rng = np.random.default_rng(1)
    Wx, Wy = rng.normal(size=(8, 3)), rng.normal(size=(8, 3))
    dcx, dcy = 6, 7
    Wx_bad, Wy_bad = Wx.copy(), Wy.copy()
    Wx_bad[dcx:] = np.nan; Wy_bad[dcy:] = np.inf
    def frag(W):
        return np.array([W.reshape(-1)[(l >> 2) * 3 + (l & 3)] if (l & 3) < 3 else 0.0 for l in range(32)])
    zero = [[0.0, 0.0]] * 32
    with np.errstate(invalid="ignore"):
        D = mma_884(frag(Wx_bad), frag(Wy_bad), zero)
    want = Wx @ Wy.T
    for lane in range(32):
        gid, tig = lane >> 2, lane & 3
        for i in range(2):
            if gid < dcx and 2 * tig + i < dcy:
                assert abs(D[lane][i] - want[gid, 2 * tig + i]) < 1e-12
