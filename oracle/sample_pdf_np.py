"""TEST INFRASTRUCTURE ONLY — bit-level numpy restatement of the reference's `sample_pdf`.

Reference: utils/run_nerf_raybased_helpers.py:283-330 (called from main.py:722-728 on CPU
tensors).  The arithmetic itself lives in PyTorch's ATen CPU kernels (third-party, pinned
torch==1.9.0 in requirements.txt:18; this image has torch 2.11.0), so this file restates
the *operation order* those kernels use, which is what a bit-exact CUDA kernel has to
reproduce:

  * ``weights + 1e-5``             one fp32 add per element (1e-5 rounded to fp32)
  * ``torch.sum(w, -1)``           ATen `vectorized_inner_sum` (SumKernel.cpp), AVX2 dispatch:
                                   8-lane vectors, 4 interleaved accumulators, then the
                                   accumulators are folded 0+=1,2,3, the scalar tail is summed
                                   first and the 8 lanes are added to it left to right.
  * ``w / sum``                    correctly rounded fp32 divide
  * ``torch.cumsum(pdf, -1)``      sequential accumulate in DOUBLE, each output rounded to fp32
  * ``torch.searchsorted(right=True)``  index = number of cdf entries <= u
  * interpolation                  separately rounded fp32 sub / div / mul / add (no FMA)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
Pinned against torch CPU and the imported reference by oracle/make_golden.py and
tests/test_oracle.py.
"""
import numpy as np

F32 = np.float32


def aten_inner_sum_f32(w: np.ndarray) -> np.ndarray:
    """Row sums of a contiguous [N, n] fp32 array in ATen-CPU (AVX2) order.  n < 512."""
    w = np.ascontiguousarray(w, dtype=F32)
    N, n = w.shape
    if n >= 512:
        raise ValueError("cascade levels of ATen's multi_row_sum (n >= 512) are not restated")
    if n < 8:
        # scalar_inner_sum: same 4-accumulator ILP scheme on scalars (rows shorter than one vector)
        a = np.zeros((4, N), dtype=F32)
        m = n // 4
        for i in range(m):
            for k in range(4):
                a[k] = (a[k] + w[:, i * 4 + k]).astype(F32)
        for i in range(m * 4, n):
            a[0] = (a[0] + w[:, i]).astype(F32)
        for k in range(1, 4):
            a[0] = (a[0] + a[k]).astype(F32)
        return a[0]
    vec = 8
    vec_size = n // vec                 # number of full 8-lane vectors
    size_ilp = vec_size // 4            # groups of 4 vectors handled by the ILP accumulators
    acc = np.zeros((4, N, vec), dtype=F32)
    for i in range(size_ilp):
        for k in range(4):
            j = (i * 4 + k) * vec
            acc[k] = (acc[k] + w[:, j:j + vec]).astype(F32)
    for i in range(size_ilp * 4, vec_size):
        j = i * vec
        acc[0] = (acc[0] + w[:, j:j + vec]).astype(F32)
    for k in range(1, 4):
        acc[0] = (acc[0] + acc[k]).astype(F32)
    final = np.zeros((N,), dtype=F32)
    for k in range(vec_size * vec, n):
        final = (final + w[:, k]).astype(F32)
    for lane in range(vec):
        final = (final + acc[0][:, lane]).astype(F32)
    return final


def sample_pdf_np(bins: np.ndarray, weights: np.ndarray, u: np.ndarray):
    """bins [N, nb], weights [N, nb-1], u [N, Ni] or [Ni] (fp32).  Returns (samples, inds, cdf).

    `u` is passed in because the reference itself builds it on the host
    (torch.linspace / torch.rand on the CPU generator, helpers:292-296).
    """
    bins = np.ascontiguousarray(bins, dtype=F32)
    w = (np.ascontiguousarray(weights, dtype=F32) + F32(1e-5)).astype(F32)       # helpers:285
    N, nb = bins.shape
    assert w.shape == (N, nb - 1)
    tot = aten_inner_sum_f32(w)                                                  # helpers:286
    pdf = (w / tot[:, None]).astype(F32)
    cdf = np.zeros((N, nb), dtype=F32)                                           # helpers:287-289
    cdf[:, 1:] = np.cumsum(pdf.astype(np.float64), axis=-1).astype(F32)
    u = np.asarray(u, dtype=F32)
    if u.ndim == 1:
        u = np.broadcast_to(u, (N, u.shape[0]))
    # searchsorted(right=True): count of entries <= u                            helpers:311-314
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)
    below = np.maximum(inds - 1, 0)                                              # helpers:315
    above = np.minimum(inds, nb - 1)                                             # helpers:316
    c0 = np.take_along_axis(cdf, below, axis=1)
    c1 = np.take_along_axis(cdf, above, axis=1)
    b0 = np.take_along_axis(bins, below, axis=1)
    b1 = np.take_along_axis(bins, above, axis=1)
    denom = (c1 - c0).astype(F32)                                                # helpers:325
    denom = np.where(denom < F32(1e-5), F32(1.0), denom).astype(F32)             # helpers:326
    t = ((u - c0).astype(F32) / denom).astype(F32)                               # helpers:327
    samples = (b0 + (t * (b1 - b0).astype(F32)).astype(F32)).astype(F32)         # helpers:328
    return samples, inds, cdf


# ---- the search-free scheme of hier_sample_det_kernel / sample_pdf_det_kernel (csrc/sample_pdf.cu), restated ----------
# Test infrastructure like the rest of oracle/: it pins the IDENTITIES the kernels rely on (for an ascending u table
# and an ascending cdf) against plain searchsorted / sort on the CPU, so a -m "not gpu" run covers the algorithm.
def inds_by_marks(cdf: np.ndarray, u: np.ndarray) -> np.ndarray:
    """searchsorted(cdf, u, right=True) = #{j : cdf_j <= u_k} without searching u per sample:
    r_j = #{k : u_k < cdf_j}; the LAST j of every run of equal r writes mark[r_j] = j + 1; inds = running max of the
    marks over k.  Needs u and cdf ascending (helpers:287-292 with det=True)."""
    N, nb = cdf.shape
    Ni = u.shape[0]
    out = np.zeros((N, Ni), dtype=np.int64)
    for n in range(N):
        r = np.searchsorted(u, cdf[n], side="left")            # #{k : u_k < cdf_j}; the kernel: ceil(cdf (Ni-1)) + fix-up
        mark = np.zeros(Ni + 1, dtype=np.int64)
        for j in range(nb):
            if j == nb - 1 or r[j] != r[j + 1]:
                mark[r[j]] = j + 1
        out[n] = np.maximum.accumulate(mark[:Ni])
    return out


def merge_by_ranks(z: np.ndarray, s: np.ndarray, inds: np.ndarray) -> np.ndarray:
    """sort(cat[z, s]) for ascending z [N, nz] and ascending samples s [N, Ni] drawn from bins `inds` (main.py:730-732):
    rank(s_k) = k + cle_k with cle_k = #{z <= s_k} = below + 1 + (z[below+1] <= s_k) (+ the kernel's verification);
    rank(z_i) = i + #{k : cle_k <= i}, again marks (last k of every run of equal cle writes k + 1) + a running max."""
    N, nz = z.shape
    Ni = s.shape[1]
    out = np.full((N, nz + Ni), np.nan, dtype=z.dtype)
    for n in range(N):
        below = np.maximum(inds[n] - 1, 0)
        zp = np.concatenate([z[n], [np.inf, np.inf, np.inf]])
        cle = below + 1 + (zp[below + 1] <= s[n])
        bad = ~(zp[below] <= s[n]) | (zp[below + 2] <= s[n])
        if bad.any():                                           # the kernel's scan fallback
            cle = np.searchsorted(z[n], s[n], side="right")
        out[n, np.arange(Ni) + cle] = s[n]
        mark = np.zeros(nz + 2, dtype=np.int64)
        for k in range(Ni):
            if k == Ni - 1 or cle[k] != cle[k + 1]:
                mark[cle[k]] = k + 1
        cnt = np.maximum.accumulate(mark[:nz])
        out[n, np.arange(nz) + cnt] = z[n]
    return out
