"""CPU oracle for the VQ bottleneck -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy, the arithmetic of the reference's
`VectorQuantizerEMA.forward` (/root/reference/src/models/vector_quantizer_ema.py:83-183)
and `VectorQuantizer.forward` (/root/reference/src/models/vector_quantizer.py:70-156),
plus the backward pass that torch autograd derives from them.  The arithmetic
itself lives in a third-party dependency of the reference, PyTorch
(requirements.txt:2, unpinned; this container has torch 2.11.0+cu128): `sum`,
`matmul`, `argmin` (first minimum), `scatter_`, `mean`, `exp`, `log`.

Parity status: the reference ships no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by `tests/golden/make_golden.py` and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` re-checks the
oracle against them on every CPU run.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this module.  The product path
(`vq-vae-speech_b200/`) never does: it fails loudly when the CUDA library is
missing.
"""
import numpy as np

NEAR_TIE_REL = 1e-6  # north_star: frames whose top-2 distance gap is below 1e-6 relative are reported, not compared


# ----------------------------------------------------------------------------------------------
# layout (vector_quantizer_ema.py:101-106)
# ----------------------------------------------------------------------------------------------
def rows_from_bdt(z):
    """(B, D, T) -> flat (N, D) rows exactly as the reference forms them.

    The reference does `inputs.permute(1, 2, 0).contiguous()` -> (D, T, B) and
    then `.view(-1, D)` (vector_quantizer_ema.py:101,106).  Row r is therefore
    flat elements [r*D, (r+1)*D) of the (D, T, B) buffer, i.e. flat element
    f = d*T*B + t*B + b <- z[b, d, t].  N = T*B.
    """
    z = np.asarray(z)
    B, D, T = z.shape
    return np.ascontiguousarray(z.transpose(1, 2, 0)).reshape(-1, D)


def bdt_from_rows(rows, B, D, T):
    """Inverse of `rows_from_bdt`: flat (N, D) -> view(D, T, B).permute(2, 0, 1) (ema.py:159,179)."""
    return np.ascontiguousarray(np.asarray(rows).reshape(D, T, B).transpose(2, 0, 1))


# ----------------------------------------------------------------------------------------------
# distance + argmin (vector_quantizer_ema.py:109-117)
# ----------------------------------------------------------------------------------------------
def distances_fp32(flat, W):
    """`(sum(x^2) + sum(e^2)) - 2 x.e^T` evaluated in fp32, same expression shape as ema.py:109-111."""
    flat = np.asarray(flat, np.float32)
    W = np.asarray(W, np.float32)
    sx = np.sum(flat * flat, axis=1, keepdims=True, dtype=np.float32)
    se = np.sum(W * W, axis=1, dtype=np.float32)
    return (sx + se) - np.float32(2.0) * (flat @ W.T)


def distances_fp64(flat, W):
    flat = np.asarray(flat, np.float64)
    W = np.asarray(W, np.float64)
    return (np.sum(flat * flat, 1, keepdims=True) + np.sum(W * W, 1)) - 2.0 * (flat @ W.T)


def assign(flat, W):
    """argmin over codes, first minimum wins (torch.argmin, ema.py:117).

    Returns (idx int64 (N,), near_tie bool (N,), gap_rel float64 (N,)).
    `near_tie[n]` marks rows whose best and second-best fp64 distances differ by
    less than NEAR_TIE_REL relative to |d_min| -- rows where two correct fp32
    implementations may legitimately disagree.
    """
    d32 = distances_fp32(flat, W)
    idx = np.argmin(d32, axis=1).astype(np.int64)
    d64 = distances_fp64(flat, W)
    K = d64.shape[1]
    if K == 1:
        return idx, np.zeros(len(idx), bool), np.full(len(idx), np.inf)
    part = np.partition(d64, 1, axis=1)
    dmin, dsec = part[:, 0], part[:, 1]
    gap_rel = (dsec - dmin) / np.maximum(np.abs(dmin), 1e-30)
    return idx, gap_rel < NEAR_TIE_REL, gap_rel


def eval_distance_tables(z, W, dtype=np.float64):
    """The three eval-only outputs of VectorQuantizer.forward (vector_quantizer.py:108-127): Euclidean distances
    torch.dist(x, y, 2) over itertools.combinations(flat_input, 2) viewed (B, -1), combinations(embedding, 2), and
    product(flat_input, embedding) viewed (B, T, K).  (The EMA class raises NameError on these lines.)"""
    z = np.asarray(z, dtype)
    W = np.asarray(W, dtype)
    B, D, T = z.shape
    flat = rows_from_bdt(z).astype(dtype)

    def pdist(a, b):
        return np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1))

    iu = np.triu_indices(flat.shape[0], 1)          # row-major upper triangle == itertools.combinations order
    enc = pdist(flat, flat)[iu].reshape(B, -1)
    ku = np.triu_indices(W.shape[0], 1)
    emb = pdist(W, W)[ku]
    fve = pdist(flat, W).reshape(B, T, -1)
    return enc, emb, fve


def one_hot(idx, K, dtype=np.float32):
    """zeros(N, K).scatter_(1, idx, 1) (ema.py:118-119)."""
    enc = np.zeros((len(idx), K), dtype)
    enc[np.arange(len(idx)), idx] = 1
    return enc


def code_stats(flat, idx, K, dtype=np.float64):
    """counts = sum(encodings, 0) (ema.py:145) and dw = encodings^T @ flat (ema.py:153)."""
    flat = np.asarray(flat, dtype)
    counts = np.bincount(idx, minlength=K).astype(dtype)
    dw = np.zeros((K, flat.shape[1]), dtype)
    np.add.at(dw, idx, flat)
    return counts, dw


def perplexity(counts, N, dtype=np.float64):
    """exp(-sum(p * log(p + 1e-10))), p = mean(encodings, 0) (ema.py:170-176)."""
    p = np.asarray(counts, dtype) / dtype(N)
    return np.exp(-np.sum(p * np.log(p + dtype(1e-10))))


# ----------------------------------------------------------------------------------------------
# EMA update (vector_quantizer_ema.py:143-156)
# ----------------------------------------------------------------------------------------------
def ema_update(cluster_size, ema_w, counts, dw, decay, eps, dtype=np.float32):
    """Returns (cluster_size', ema_w', W').  Laplace smoothing exactly as ema.py:147-151."""
    f = dtype
    cs = np.asarray(cluster_size, f) * f(decay) + f(1 - decay) * np.asarray(counts, f)
    n = np.sum(cs, dtype=f)
    K = cs.shape[0]
    cs = (cs + f(eps)) / (n + f(K) * f(eps)) * n
    ew = np.asarray(ema_w, f) * f(decay) + f(1 - decay) * np.asarray(dw, f)
    W = ew / cs[:, None]
    return cs, ew, W


# ----------------------------------------------------------------------------------------------
# full forward / backward
# ----------------------------------------------------------------------------------------------
def vq_forward(z, W, commitment_cost, ema=None, training=True, dtype=np.float64, stats_allreduce=None):
    """Forward of VectorQuantizer (ema=None) or VectorQuantizerEMA (ema=dict).

    z: (B, D, T).  W: (K, D) codebook (`_embedding.weight`).
    ema: None, or {'cluster_size': (K,), 'ema_w': (K, D), 'decay': float, 'eps': float}.
    Index search is always fp32 (that is what the reference computes); every
    other quantity is evaluated in `dtype` (float64 = high-accuracy truth used
    for the 1e-5 relative parity bound, float32 = like-for-like).

    stats_allreduce: optional callable (counts, dw, n_rows) -> (counts, dw, n_rows_total) that sums the per-code
    statistics over data-parallel shards: the EMA update and the perplexity then use the global
    statistics, everything else (indices, quantisation, loss) stays per shard.

    Returns a dict: idx, near_tie, gap_rel, counts, dw, W_used (codebook used
    for quantisation: updated one in EMA training mode, ema.py:156-159),
    cluster_size / ema_w (new EMA state), quantized (B, D, T), e_latent,
    q_latent (non-EMA), vq_loss, perplexity, encodings (B, T, K) view,
    distances (B, T, K) view (fp32).
    """
    z = np.asarray(z)
    B, D, T = z.shape
    K = W.shape[0]
    flat32 = rows_from_bdt(z.astype(np.float32))
    N = flat32.shape[0]
    idx, near, gap = assign(flat32, W)
    flat = flat32.astype(dtype)
    counts, dw = code_stats(flat, idx, K, dtype)
    out = dict(idx=idx, near_tie=near, gap_rel=gap, counts=counts, dw=dw, N=N)
    Wq = np.asarray(W, dtype)
    g_counts, g_dw, n_total = counts, dw, N
    if stats_allreduce is not None and ema is not None and training:
        g_counts, g_dw, n_total = stats_allreduce(counts, dw, N)
    if ema is not None and training:
        cs, ew, Wq = ema_update(ema['cluster_size'], ema['ema_w'], g_counts, g_dw, ema['decay'], ema['eps'], dtype)
        out.update(cluster_size=cs, ema_w=ew)
    out['W_used'] = Wq
    q_rows = Wq[idx]                                   # matmul(encodings, W) (ema.py:159)
    diff = q_rows - flat
    e_latent = np.mean(diff * diff, dtype=dtype)       # ema.py:165
    out['e_latent'] = e_latent
    if ema is None:
        out['q_latent'] = e_latent                     # same value, different gradient (vector_quantizer.py:136-137)
        out['vq_loss'] = e_latent + dtype(commitment_cost) * e_latent
    else:
        out['vq_loss'] = dtype(commitment_cost) * e_latent
    ste_rows = flat + (q_rows - flat)                  # inputs + (quantized - inputs).detach() (ema.py:169)
    out['quantized'] = bdt_from_rows(ste_rows, B, D, T)
    out['q_rows'] = q_rows
    out['perplexity'] = perplexity(g_counts, n_total, dtype)
    out['encodings'] = one_hot(idx, K).reshape(B, T, K)          # .view(batch_size, time, -1) of an (N, K) buffer
    out['distances'] = distances_fp32(flat32, W).reshape(B, T, K)
    return out


def vq_backward(z, fwd, commitment_cost, g_quantized, g_loss=1.0, ema=False, dtype=np.float64):
    """What autograd derives from ema.py:165-169 / vector_quantizer.py:136-141.

    grad_z = g_quantized + g_loss * beta * 2 (x - q) / (N*D)          (both variants)
    grad_E[k] = g_loss * (2 / (N*D)) * (count_k * E_k - dw_k)         (non-EMA only; EMA codebook gets none)
    """
    z = np.asarray(z, dtype)
    B, D, T = z.shape
    N = fwd['N']
    numel = dtype(N * D)
    flat = rows_from_bdt(z)
    diff_rows = flat - fwd['q_rows']
    gz_rows = dtype(g_loss) * dtype(commitment_cost) * dtype(2.0) * diff_rows / numel
    grad_z = np.asarray(g_quantized, dtype) + bdt_from_rows(gz_rows, B, D, T)
    grad_E = None
    if not ema:
        grad_E = dtype(g_loss) * (dtype(2.0) / numel) * (fwd['counts'][:, None] * fwd['W_used'] - fwd['dw'])
    return grad_z, grad_E
