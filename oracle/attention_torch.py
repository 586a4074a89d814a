"""TEST INFRASTRUCTURE -- gradients of the temporal-attention loss (original_attention.py:88-152) by torch.autograd over a
float64 CPU restatement of the same forward (statement for statement the one in oracle/attention_numpy.py, which the CPU
tests hold it equal to), plus the optimiser of train() (:427-435): tf.clip_by_global_norm(gradients, 10) -- with the Wemb
gradient entering the norm as TF's un-deduplicated IndexedSlices (one slice per embedding_lookup row) -- and
tf.train.AdamOptimizer's update (lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), epsilon outside the square root)."""
import numpy as np
import torch

from .attention_numpy import LSTM3_B, LSTM3_W, REG_FRAMES


def loss_and_grads(p, video, caption, caption_mask, drop_mult=None, beta=10.0, m=0.5):
    """-> (loss, regulariser part, {name: gradient ndarray}, Wemb IndexedSlices square norm)."""
    P = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in p.items()}
    video = torch.tensor(np.asarray(video, dtype=np.float64))
    cap = torch.tensor(np.asarray(caption, dtype=np.int64))
    mask = torch.tensor(np.asarray(caption_mask, dtype=np.float64))
    B, n, D = video.shape
    H = P['embed_att_Wa'].shape[0]
    Tc = cap.shape[1]
    emb = (video.reshape(-1, D) @ P['encode_image_W'] + P['encode_image_b']).reshape(B, n, H).permute(1, 0, 2)     # n x b x h
    part = emb @ P['embed_att_Ua'] + P['embed_att_ba']
    c = torch.zeros(B, H, dtype=torch.float64); h = torch.zeros(B, H, dtype=torch.float64)
    h_prev = torch.zeros(B, H, dtype=torch.float64)
    current_embed = torch.zeros(B, H, dtype=torch.float64)
    lookups = []
    loss_caption = 0.0; reg_total = 0.0
    for i in range(Tc):
        e = (torch.tanh(h_prev @ P['embed_att_Wa'] + part) @ P['embed_att_w'])[:, :, 0]
        e_hat_exp = torch.exp(e)
        denomin = e_hat_exp.sum(0)
        denomin = denomin + (denomin == 0).to(torch.float64)
        alphas = e_hat_exp / denomin
        atten = (alphas[:, :, None] * emb).sum(0)
        z = torch.cat([atten, current_embed, h], 1) @ P[LSTM3_W] + P[LSTM3_B]
        ii, jj, ff, oo = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
        c = c * torch.sigmoid(ff + 1.0) + torch.sigmoid(ii) * torch.tanh(jj)
        h = torch.tanh(c) * torch.sigmoid(oo)
        out1 = h if drop_mult is None else h * torch.tensor(np.asarray(drop_mult[i], dtype=np.float64))
        out2 = torch.tanh(torch.cat([out1, atten, current_embed], 1) @ P['embed_nn_Wp'] + P['embed_nn_bp'])
        h_prev = out1
        current_embed = P['Wemb'][cap[:, i]]
        current_embed.retain_grad()
        lookups.append(current_embed)
        logit_words = out2 @ P['embed_word_W'] + P['embed_word_b']
        cross_entropy = torch.logsumexp(logit_words, 1) - logit_words[torch.arange(B), cap[:, i]]
        regularizer = beta * torch.clamp(m - alphas.t()[:, 0:REG_FRAMES].sum(1), min=0.0) * mask[:, i]
        loss_caption = loss_caption + (cross_entropy * mask[:, i] + regularizer).sum()
        reg_total = reg_total + regularizer.sum()
    loss = loss_caption / mask.sum()
    loss.backward()
    grads = {k: (v.grad.numpy().copy() if v.grad is not None else np.zeros(v.shape)) for k, v in P.items()}
    slice_sq = float(sum((t.grad ** 2).sum().item() for t in lookups if t.grad is not None))
    return float(loss.item()), float((reg_total / mask.sum()).item()), grads, slice_sq


def clip_and_adam(p, grads, slice_sq, state, lr, clip_norm=10.0, b1=0.9, b2=0.999, eps=1e-8):
    """One train_op: global norm with the Wemb IndexedSlices norm, clip, Adam.  state = {'t', 'm', 'v'} (mutated).
    -> (new params, global norm)."""
    sq = sum(float((g.astype(np.float64) ** 2).sum()) for k, g in grads.items() if k != 'Wemb') + slice_sq
    gn = np.sqrt(sq)
    scale = clip_norm * min(1.0 / gn, 1.0 / clip_norm) if gn > 0 else 1.0
    state['t'] += 1
    lr_t = lr * np.sqrt(1.0 - b2 ** state['t']) / (1.0 - b1 ** state['t'])
    out = {}
    for k in p:
        g = grads[k] * scale
        state['m'][k] = b1 * state['m'].get(k, 0.0) + (1 - b1) * g
        state['v'][k] = b2 * state['v'].get(k, 0.0) + (1 - b2) * g * g
        out[k] = p[k] - lr_t * state['m'][k] / (np.sqrt(state['v'][k]) + eps)
    return out, gn
