"""CPU oracle for the Conv1d encoder / residual stack / jitter / decoder / training step.

TEST INFRASTRUCTURE ONLY (see oracle/vq_oracle.py for the rules on who may import it).

Plain-numpy restatement, forward AND hand-derived backward, of

  * ConvolutionalEncoder.forward      /root/reference/src/models/convolutional_encoder.py:118-146
  * Residual / ResidualStack          /root/reference/src/modules/residual.py:31-70, residual_stack.py:34-46
  * Jitter.forward                    /root/reference/src/modules/jitter.py:47-70
  * DeconvolutionalDecoder.forward    /root/reference/src/models/deconvolutional_decoder.py:100-137
  * ConvolutionalVQVAE.forward        /root/reference/src/models/convolutional_vq_vae.py:117-139
  * ConvolutionalTrainer.iterate      /root/reference/src/experiments/convolutional_trainer.py:44-74
    (MSELoss + vq_loss, backward, Adam(lr, amsgrad=True))

The arithmetic of `nn.Conv1d`, `nn.ConvTranspose1d`, `nn.Upsample(scale_factor=2)` (nearest),
`F.relu`, `nn.MSELoss` and `optim.Adam` lives in PyTorch (requirements.txt:2, unpinned;
torch 2.11.0+cu128 here); their published definitions are restated below.  Pinned against
the reference itself by tests/golden/make_golden.py -> tests/golden/model_*.npz.
"""
import numpy as np

from . import vq_oracle as vqo


# ----------------------------------------------------------------------------------------------
# primitive ops: nn.Conv1d / nn.ConvTranspose1d (cross-correlation definition in the torch docs)
# ----------------------------------------------------------------------------------------------
def conv1d_fwd(x, w, b=None, stride=1, pad=0):
    """y[n,o,l] = b[o] + sum_{c,j} w[o,c,j] * xpad[n,c,l*stride+j]."""
    B, Ci, L = x.shape
    Co, _, k = w.shape
    Lo = (L + 2 * pad - k) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad)))
    y = np.zeros((B, Co, Lo), x.dtype)
    for j in range(k):
        y += np.einsum('oc,bcl->bol', w[:, :, j], xp[:, :, j:j + (Lo - 1) * stride + 1:stride])
    if b is not None:
        y += b[None, :, None]
    return y


def conv1d_dgrad(dy, w, L, stride=1, pad=0):
    """dx[n,c,i] = sum_{o,j} w[o,c,j] * dy[n,o,l] with l*stride + j - pad = i."""
    B, Co, Lo = dy.shape
    _, Ci, k = w.shape
    dxp = np.zeros((B, Ci, L + 2 * pad), dy.dtype)
    for j in range(k):
        dxp[:, :, j:j + (Lo - 1) * stride + 1:stride] += np.einsum('oc,bol->bcl', w[:, :, j], dy)
    return dxp[:, :, pad:pad + L]


def conv1d_wgrad(dy, x, k, stride=1, pad=0, bias=True):
    """dw[o,c,j] = sum_{n,l} dy[n,o,l] * xpad[n,c,l*stride+j];  db[o] = sum_{n,l} dy[n,o,l]."""
    B, Co, Lo = dy.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad)))
    dw = np.zeros((Co, x.shape[1], k), dy.dtype)
    for j in range(k):
        dw[:, :, j] = np.einsum('bol,bcl->oc', dy, xp[:, :, j:j + (Lo - 1) * stride + 1:stride])
    db = dy.sum(axis=(0, 2)) if bias else None
    return dw, db


def convT1d_fwd(x, w, b=None, pad=0):
    """nn.ConvTranspose1d, stride 1: y[n,o,i] = b[o] + sum_{c,j} x[n,c,l] * w[c,o,j] with i = l + j - pad.

    Equals conv1d with w'[o,c,j] = w[c,o,k-1-j] and padding k-1-pad."""
    k = w.shape[2]
    wf = np.ascontiguousarray(w[:, :, ::-1].transpose(1, 0, 2))
    return conv1d_fwd(x, wf, b, 1, k - 1 - pad)


def convT1d_dgrad(dy, w, L, pad=0):
    """dx[n,c,l] = sum_{o,j} dy[n,o,l+j-pad] * w[c,o,j]  (a plain conv1d of dy with weight w, padding `pad`)."""
    return conv1d_fwd(dy, w, None, 1, pad)


def convT1d_wgrad(dy, x, k, pad=0):
    """dw[c,o,j] = sum_{n,l} x[n,c,l] * dy[n,o,l+j-pad];  db[o] = sum dy."""
    B, Ci, L = x.shape
    Co = dy.shape[1]
    dyp = np.pad(dy, ((0, 0), (0, 0), (pad, pad)))
    dw = np.zeros((Ci, Co, k), dy.dtype)
    for j in range(k):
        dw[:, :, j] = np.einsum('bcl,bol->co', x, dyp[:, :, j:j + L])
    return dw, dy.sum(axis=(0, 2))


def relu(x):
    return np.maximum(x, 0)


def upsample2(x):
    """nn.Upsample(scale_factor=2), mode nearest: out[..., i] = in[..., i // 2]."""
    return np.repeat(x, 2, axis=2)


def upsample2_bwd(g):
    return g[:, :, 0::2] + g[:, :, 1::2]


# ----------------------------------------------------------------------------------------------
# Residual stack (shared weights; in-place ReLU)
# ----------------------------------------------------------------------------------------------
def resstack_fwd(x, w1, w2, n_layers):
    """x_{i+1} = relu(x_i) + conv2(relu(conv1(relu(x_i)))); returns relu(x_n) and the cache.

    The block starts with nn.ReLU(inplace=True) (residual.py:36) so the skip term
    sees relu(x_i), not x_i; the SAME Residual instance is applied n_layers times
    (residual_stack.py:40-41)."""
    cache = []
    for _ in range(n_layers):
        a = relu(x)
        h = relu(conv1d_fwd(a, w1, None, 1, 1))
        y = a + conv1d_fwd(h, w2, None, 1, 0)
        cache.append((x, a, h))
        x = y
    return relu(x), (cache, x)


def resstack_bwd(g, w1, w2, cache):
    layers, x_last = cache
    g = g * (x_last > 0)
    dw1 = np.zeros_like(w1)
    dw2 = np.zeros_like(w2)
    for (x, a, h) in reversed(layers):
        d2, _ = conv1d_wgrad(g, h, 1, 1, 0, bias=False)
        gh = conv1d_dgrad(g, w2, h.shape[2], 1, 0) * (h > 0)
        d1, _ = conv1d_wgrad(gh, a, 3, 1, 1, bias=False)
        ga = g + conv1d_dgrad(gh, w1, a.shape[2], 1, 1)
        g = ga * (x > 0)
        dw1 += d1
        dw2 += d2
    return g, dw1, dw2


# ----------------------------------------------------------------------------------------------
# Jitter (jitter.py:47-70): host RNG plan, consumed in exactly the reference's order
# ----------------------------------------------------------------------------------------------
def jitter_plan(length, probability, rng=np.random):
    """src[t] = column of the ORIGINAL tensor that ends up at t.  One choice([1,0],p) per t and,
    only for a replaced interior t, one choice([-1,1]) (jitter.py:55-67)."""
    src = np.arange(length, dtype=np.int64)
    for i in range(length):
        replace = [True, False][rng.choice([1, 0], p=[probability, 1 - probability])]
        if replace:
            if i == 0:
                nb = i + 1
            elif i == length - 1:
                nb = i - 1
            else:
                nb = i + rng.choice([-1, 1], p=[0.5, 0.5])
            src[i] = nb
    return src


# ----------------------------------------------------------------------------------------------
# full model
# ----------------------------------------------------------------------------------------------
ENC = '_encoder.'
DEC = '_decoder.'
RS1 = '_residual_stack._layers.0._block.1.weight'
RS2 = '_residual_stack._layers.0._block.3.weight'


def wn_expand(p, dtype=np.float64):
    """use_kaiming_normal: every conv is wrapped in nn.utils.weight_norm (conv1d_builder.py:41-43,
    conv_transpose1d_builder.py:41-43, residual.py:45-47,57-59), i.e. weight = g * v / ||v|| with the norm over all
    dims but 0 (torch._weight_norm, dim = 0).  Returns p plus the derived '<layer>.weight' entries."""
    if not any(k.endswith('.weight_v') for k in p):
        return p
    q = dict(p)
    for k in p:
        if k.endswith('.weight_v'):
            v = np.asarray(p[k], dtype)
            g = np.asarray(p[k[:-2] + '_g'], dtype)
            n = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(1)).reshape((-1,) + (1,) * (v.ndim - 1))
            q[k[:-2]] = v * (g / n)
    return q


def wn_fold_grads(p, grads, dtype=np.float64):
    """Autograd of the reparametrisation: d/dg = <dW, v> / ||v||,  d/dv = g / ||v|| * (dW - v <dW, v> / ||v||^2)."""
    for k in p:
        if k.endswith('.weight_v') and k[:-2] in grads:
            v = np.asarray(p[k], dtype)
            g = np.asarray(p[k[:-2] + '_g'], dtype)
            dW = grads.pop(k[:-2])
            shp = (-1,) + (1,) * (v.ndim - 1)
            n = np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(1)).reshape(shp)
            dot = (dW * v).reshape(v.shape[0], -1).sum(1).reshape(shp)
            grads[k[:-2] + '_g'] = dot / n
            grads[k] = (g / n) * (dW - v * dot / (n * n))
    return grads


def trainable_names(params, ema):
    """Parameters Adam actually updates.  In EMA mode `_vq._embedding.weight` / `_vq._ema_w` are
    re-created every step (ema.py:154,156) so the optimizer's references never see a gradient."""
    names = []
    for k in params:
        if k.endswith('_ema_cluster_size'):
            continue
        if '_layers.1.' in k:          # same tensor as _layers.0 (residual_stack.py:40-41)
            continue
        if ema and k.startswith('_vq.'):
            continue
        names.append(k)
    return names


def model_forward(p, x_btf, cfg, jitter_src=None, training=True, dtype=np.float64):
    """ConvolutionalVQVAE.forward.  x_btf: (B, T, F).  p: dict name -> array (state_dict names).
    cfg: commitment_cost, decay, epsilon, num_residual_layers.  Returns (outputs dict, cache)."""
    c = {}
    p = wn_expand(p, dtype)
    g = lambda n: np.asarray(p[n], dtype)
    nl = cfg['num_residual_layers']
    x = np.ascontiguousarray(np.asarray(x_btf, dtype).transpose(0, 2, 1))      # vq_vae.py:118
    c['x'] = x
    c['p1'] = conv1d_fwd(x, g(ENC + '_conv_1.weight'), g(ENC + '_conv_1.bias'), 1, 1)
    a1 = relu(c['p1'])
    c['a1'] = a1
    c['p2'] = conv1d_fwd(a1, g(ENC + '_conv_2.weight'), g(ENC + '_conv_2.bias'), 1, 1)
    h2 = relu(c['p2']) + a1
    c['h2'] = h2
    c['p3'] = conv1d_fwd(h2, g(ENC + '_conv_3.weight'), g(ENC + '_conv_3.bias'), 2, 2)
    a3 = relu(c['p3'])
    c['a3'] = a3
    c['p4'] = conv1d_fwd(a3, g(ENC + '_conv_4.weight'), g(ENC + '_conv_4.bias'), 1, 1)
    h4 = relu(c['p4']) + a3
    c['h4'] = h4
    c['p5'] = conv1d_fwd(h4, g(ENC + '_conv_5.weight'), g(ENC + '_conv_5.bias'), 1, 1)
    h5 = relu(c['p5']) + h4
    c['h5'] = h5
    r, c['enc_rs'] = resstack_fwd(h5, g(ENC + RS1), g(ENC + RS2), nl)
    enc_out = r + h5
    c['enc_out'] = enc_out
    z = conv1d_fwd(enc_out, g('_pre_vq_conv.weight'), g('_pre_vq_conv.bias'), 1, 1)
    c['z'] = z

    ema = None
    if cfg['decay'] > 0.0:
        ema = dict(cluster_size=g('_vq._ema_cluster_size'), ema_w=g('_vq._ema_w'),
                   decay=cfg['decay'], eps=cfg.get('epsilon', 1e-5))
    vq = vqo.vq_forward(z, np.asarray(p['_vq._embedding.weight'], np.float32), cfg['commitment_cost'],
                        ema=ema, training=training, dtype=dtype, stats_allreduce=cfg.get('stats_allreduce'))
    c['vq'] = vq
    q = vq['quantized']
    if jitter_src is not None:
        q = q[:, :, jitter_src]
    c['jitter_src'] = jitter_src
    c['n_spk'] = 0
    if cfg.get('speaker_features') is not None:
        # deconvolutional_decoder.py:108-111: (B, 40) speaker features repeated over time, concatenated on the channel axis
        sf = np.asarray(cfg['speaker_features'], dtype)
        c['n_spk'] = sf.shape[1]
        q = np.concatenate([q, np.repeat(sf[:, :, None], q.shape[2], axis=2)], axis=1)
    c['dec_in'] = q
    d1 = conv1d_fwd(q, g(DEC + '_conv_1.weight'), g(DEC + '_conv_1.bias'), 1, 1)
    u = upsample2(d1)
    s, c['dec_rs'] = resstack_fwd(u, g(DEC + RS1), g(DEC + RS2), nl)
    c['s'] = s
    c['q1'] = convT1d_fwd(s, g(DEC + '_conv_trans_1.weight'), g(DEC + '_conv_trans_1.bias'), 1)
    t1 = relu(c['q1'])
    c['t1'] = t1
    c['q2'] = convT1d_fwd(t1, g(DEC + '_conv_trans_2.weight'), g(DEC + '_conv_trans_2.bias'), 0)
    t2 = relu(c['q2'])
    c['t2'] = t2
    t3 = convT1d_fwd(t2, g(DEC + '_conv_trans_3.weight'), g(DEC + '_conv_trans_3.bias'), 0)
    c['Lout'] = t3.shape[2]
    T = x.shape[2]
    recon = t3[:, :, :T]                                  # vq_vae.py:133-137 (drop the last Lout - T steps)
    out = dict(reconstructed_x=recon, vq_loss=vq['vq_loss'], perplexity=vq['perplexity'],
               encoding_indices=vq['idx'].reshape(-1, 1), z=z, quantized=vq['quantized'])
    return out, c


def model_backward(p, c, out, target_bft, cfg, dtype=np.float64):
    """d(loss)/d(params) for loss = vq_loss + mean((recon - target)^2)  (trainer.py:54-63)."""
    p_raw, p = p, wn_expand(p, dtype)
    g = lambda n: np.asarray(p[n], dtype)
    recon = out['reconstructed_x']
    target = np.asarray(target_bft, dtype)
    diff = recon - target
    recon_loss = np.mean(diff * diff)
    grads = {}
    B, Fo, T = recon.shape
    gt3 = np.zeros((B, Fo, c['Lout']), dtype)
    gt3[:, :, :T] = 2.0 * diff / diff.size
    # decoder
    w = g(DEC + '_conv_trans_3.weight')
    grads[DEC + '_conv_trans_3.weight'], grads[DEC + '_conv_trans_3.bias'] = convT1d_wgrad(gt3, c['t2'], w.shape[2], 0)
    gq2 = convT1d_dgrad(gt3, w, c['t2'].shape[2], 0) * (c['q2'] > 0)
    w = g(DEC + '_conv_trans_2.weight')
    grads[DEC + '_conv_trans_2.weight'], grads[DEC + '_conv_trans_2.bias'] = convT1d_wgrad(gq2, c['t1'], w.shape[2], 0)
    gq1 = convT1d_dgrad(gq2, w, c['t1'].shape[2], 0) * (c['q1'] > 0)
    w = g(DEC + '_conv_trans_1.weight')
    grads[DEC + '_conv_trans_1.weight'], grads[DEC + '_conv_trans_1.bias'] = convT1d_wgrad(gq1, c['s'], w.shape[2], 1)
    gs = convT1d_dgrad(gq1, w, c['s'].shape[2], 1)
    gu, grads[DEC + RS1], grads[DEC + RS2] = resstack_bwd(gs, g(DEC + RS1), g(DEC + RS2), c['dec_rs'])
    gd1 = upsample2_bwd(gu)
    w = g(DEC + '_conv_1.weight')
    grads[DEC + '_conv_1.weight'], grads[DEC + '_conv_1.bias'] = conv1d_wgrad(gd1, c['dec_in'], 3, 1, 1)
    gq = conv1d_dgrad(gd1, w, c['dec_in'].shape[2], 1, 1)
    if c['n_spk']:
        gq = gq[:, :gq.shape[1] - c['n_spk']]              # the speaker features are constants of the step
    if c['jitter_src'] is not None:
        # in-place column copies from a detached clone: replaced columns get zero gradient (jitter.py:49,68)
        keep = (c['jitter_src'] == np.arange(len(c['jitter_src'])))
        gq = gq * keep[None, None, :]
    # VQ
    ema = cfg['decay'] > 0.0
    gz, gE = vqo.vq_backward(c['z'], c['vq'], cfg['commitment_cost'], gq, 1.0, ema=ema, dtype=dtype)
    if not ema:
        grads['_vq._embedding.weight'] = gE
    # pre-VQ conv
    w = g('_pre_vq_conv.weight')
    grads['_pre_vq_conv.weight'], grads['_pre_vq_conv.bias'] = conv1d_wgrad(gz, c['enc_out'], 3, 1, 1)
    ge = conv1d_dgrad(gz, w, c['enc_out'].shape[2], 1, 1)
    # encoder
    gr, grads[ENC + RS1], grads[ENC + RS2] = resstack_bwd(ge, g(ENC + RS1), g(ENC + RS2), c['enc_rs'])
    gh5 = ge + gr
    gp5 = gh5 * (c['p5'] > 0)
    w = g(ENC + '_conv_5.weight')
    grads[ENC + '_conv_5.weight'], grads[ENC + '_conv_5.bias'] = conv1d_wgrad(gp5, c['h4'], 3, 1, 1)
    gh4 = gh5 + conv1d_dgrad(gp5, w, c['h4'].shape[2], 1, 1)
    gp4 = gh4 * (c['p4'] > 0)
    w = g(ENC + '_conv_4.weight')
    grads[ENC + '_conv_4.weight'], grads[ENC + '_conv_4.bias'] = conv1d_wgrad(gp4, c['a3'], 3, 1, 1)
    ga3 = gh4 + conv1d_dgrad(gp4, w, c['a3'].shape[2], 1, 1)
    gp3 = ga3 * (c['p3'] > 0)
    w = g(ENC + '_conv_3.weight')
    grads[ENC + '_conv_3.weight'], grads[ENC + '_conv_3.bias'] = conv1d_wgrad(gp3, c['h2'], 4, 2, 2)
    gh2 = conv1d_dgrad(gp3, w, c['h2'].shape[2], 2, 2)
    gp2 = gh2 * (c['p2'] > 0)
    w = g(ENC + '_conv_2.weight')
    grads[ENC + '_conv_2.weight'], grads[ENC + '_conv_2.bias'] = conv1d_wgrad(gp2, c['a1'], 3, 1, 1)
    ga1 = gh2 + conv1d_dgrad(gp2, w, c['a1'].shape[2], 1, 1)
    gp1 = ga1 * (c['p1'] > 0)
    grads[ENC + '_conv_1.weight'], grads[ENC + '_conv_1.bias'] = conv1d_wgrad(gp1, c['x'], 3, 1, 1)
    return wn_fold_grads(p_raw, grads, dtype), recon_loss


def amsgrad_step(param, grad, m, v, vmax, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam(amsgrad=True), weight_decay 0, single-tensor formulation (torch/optim/adam.py)."""
    dt = param.dtype.type
    m = m * dt(beta1) + dt(1 - beta1) * grad
    v = v * dt(beta2) + dt(1 - beta2) * grad * grad
    vmax = np.maximum(vmax, v)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(vmax) / dt(np.sqrt(bc2)) + dt(eps)
    param = param - dt(lr / bc1) * (m / denom)
    return param, m, v, vmax


def train_step(p, opt, x_btf, cfg, jitter_src=None, dtype=np.float64):
    """One ConvolutionalTrainer.iterate (trainer.py:44-74).  p and opt are updated IN PLACE.
    opt: {'step': int, 'm': {}, 'v': {}, 'vmax': {}}.  target = the input features."""
    out, c = model_forward(p, x_btf, cfg, jitter_src, True, dtype)
    target = np.asarray(x_btf, dtype).transpose(0, 2, 1)
    grads, recon_loss = model_backward(p, c, out, target, cfg, dtype)
    ema = cfg['decay'] > 0.0
    opt['step'] += 1
    for n in trainable_names(p, ema):
        if n not in opt['m']:
            z = np.zeros_like(np.asarray(p[n], dtype))
            opt['m'][n], opt['v'][n], opt['vmax'][n] = z, z.copy(), z.copy()
        newp, opt['m'][n], opt['v'][n], opt['vmax'][n] = amsgrad_step(
            np.asarray(p[n], dtype), grads[n], opt['m'][n], opt['v'][n], opt['vmax'][n],
            opt['step'], cfg['learning_rate'])
        p[n] = newp
    for n in list(p):
        if '_layers.1.' in n:
            p[n] = p[n.replace('_layers.1.', '_layers.0.')]
    if ema:
        vq = c['vq']
        p['_vq._ema_cluster_size'] = vq['cluster_size']
        p['_vq._ema_w'] = vq['ema_w']
        p['_vq._embedding.weight'] = vq['W_used']
    return dict(loss=out['vq_loss'] + recon_loss, reconstruction_loss=recon_loss, vq_loss=out['vq_loss'],
                perplexity=out['perplexity'], encoding_indices=out['encoding_indices'],
                reconstructed_x=out['reconstructed_x'], grads=grads, near_tie=c['vq']['near_tie'])
