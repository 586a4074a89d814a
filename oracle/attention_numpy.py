"""TEST INFRASTRUCTURE -- NumPy restatement of the temporal-attention caption model of original_attention.py
(BASELINE config 3): single LSTM3 decoder with additive attention over the frame embeddings, a tanh MLP
head and the hinge regulariser on the attention weights of the first 8 frames.

    __init__        original_attention.py:54-86   (variables)
    build_model     :88-152   (teacher-forced loss with DropoutWrapper on the LSTM3 output)
    build_generator :155-199  (greedy decode)     build_sampler :201-251 (greedy decode + saved alphas)

TF-1.1 library semantics as in oracle/s2vt_numpy.py: BasicLSTMCell gate order i, j, f, o with forget bias 1 on the kernel
[inputs ; h] -> 4H; DropoutWrapper(output_keep_prob) scales the OUTPUT only (the recurrent state stays un-dropped);
softmax_cross_entropy_with_logits on one-hot labels.  The attention softmax is the reference's literal exp / sum (no
max subtraction) with the `denominator == 0 -> +1` guard (:117-121).
PARITY UNPINNED (the reference has no tests); pinned against regressions by tests/golden/attention_golden.npz.
"""
import numpy as np

LSTM3_W = 's2vt/LSTM3/basic_lstm_cell/weights'
LSTM3_B = 's2vt/LSTM3/basic_lstm_cell/biases'
REG_FRAMES = 8          # alphas_1 = temp_alphas[:, 0:8]  (:123)


def init_params(D, H, V, seed=16, dtype=np.float64):
    """Variables of :64-86: U(-0.1, 0.1) matrices, zero biases, glorot-uniform LSTM3 kernel [3H, 4H]."""
    rng = np.random.RandomState(seed)
    u = lambda *s: rng.uniform(-0.1, 0.1, size=s).astype(dtype)
    lim = np.sqrt(6.0 / (3 * H + 4 * H))
    return {
        'Wemb': u(V, H), 'encode_image_W': u(D, H), 'encode_image_b': np.zeros(H, dtype),
        'embed_att_w': u(H, 1), 'embed_att_Wa': u(H, H), 'embed_att_Ua': u(H, H), 'embed_att_ba': np.zeros(H, dtype),
        'embed_word_W': u(H, V), 'embed_word_b': np.zeros(V, dtype),
        'embed_nn_Wp': u(3 * H, H), 'embed_nn_bp': np.zeros(H, dtype),
        LSTM3_W: rng.uniform(-lim, lim, size=(3 * H, 4 * H)).astype(dtype), LSTM3_B: np.zeros(4 * H, dtype),
    }


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _frames(p, video):
    """:95-98, 107: image_emb [n, b, h] and image_part = image_emb . Ua + ba."""
    B, n, D = video.shape
    emb = video.reshape(-1, D) @ p['encode_image_W'] + p['encode_image_b']
    emb = emb.reshape(B, n, -1).transpose(1, 0, 2)
    part = emb @ p['embed_att_Ua'] + p['embed_att_ba']
    return emb, part


def _attend(p, h_prev, emb, part):
    """:113-128: alphas [n, b] and the attended frame vector [b, h]."""
    e = np.tanh(h_prev @ p['embed_att_Wa'] + part)                 # n x b x h
    e = (e @ p['embed_att_w'])[:, :, 0]                            # n x b
    e_hat_exp = np.exp(e)
    denomin = e_hat_exp.sum(0)
    denomin = denomin + (denomin == 0).astype(e.dtype)
    alphas = e_hat_exp / denomin
    atten = (alphas[:, :, None] * emb).sum(0)
    return alphas, atten


def _lstm3(p, x, state):
    """BasicLSTMCell(state_is_tuple=False): state = [c, h]; returns (h', [c', h'])."""
    H = state.shape[1] // 2
    c, h = state[:, :H], state[:, H:]
    z = np.concatenate([x, h], 1) @ p[LSTM3_W] + p[LSTM3_B]
    i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
    c2 = c * _sigmoid(f + 1.0) + _sigmoid(i) * np.tanh(j)
    h2 = np.tanh(c2) * _sigmoid(o)
    return h2, np.concatenate([c2, h2], 1)


def build_model_loss(p, video, caption, caption_mask, drop_mult=None, beta=10.0, m=0.5, return_logits=False):
    """build_model :88-152.  drop_mult: [T_c, B, H] multipliers (0 or 1/keep) of the DropoutWrapper, None = keep 1.
    Returns (loss, regulariser part of the loss[, logits [T_c, B, V]])."""
    B = video.shape[0]
    H = p['embed_att_Wa'].shape[0]
    Tc = caption.shape[1]
    dt = p['Wemb'].dtype
    emb, part = _frames(p, video.astype(dt))
    state = np.zeros((B, 2 * H), dt)
    h_prev = np.zeros((B, H), dt)
    current_embed = np.zeros((B, H), dt)
    loss_caption, reg_total, logits_all = 0.0, 0.0, []
    for i in range(Tc):
        alphas, atten = _attend(p, h_prev, emb, part)
        out1, state = _lstm3(p, np.concatenate([atten, current_embed], 1), state)
        if drop_mult is not None:
            out1 = out1 * drop_mult[i].astype(dt)
        out2 = np.tanh(np.concatenate([out1, atten, current_embed], 1) @ p['embed_nn_Wp'] + p['embed_nn_bp'])
        h_prev = out1
        current_embed = p['Wemb'][caption[:, i]]
        logit_words = out2 @ p['embed_word_W'] + p['embed_word_b']
        mx = logit_words.max(1, keepdims=True)
        lse = mx[:, 0] + np.log(np.exp(logit_words - mx).sum(1))
        cross_entropy = lse - logit_words[np.arange(B), caption[:, i]]
        regularizer = beta * np.maximum(0.0, m - alphas.T[:, 0:REG_FRAMES].sum(1)) * caption_mask[:, i]
        loss_caption += (cross_entropy * caption_mask[:, i] + regularizer).sum()
        reg_total += regularizer.sum()
        if return_logits:
            logits_all.append(logit_words)
    norm = caption_mask.sum()
    out = (loss_caption / norm, reg_total / norm)
    return out + (np.stack(logits_all),) if return_logits else out


def build_sampler(p, video, n_caption_lstm_steps=35, return_logits=False):
    """build_sampler :201-251 (= build_generator :155-199 plus the alphas): greedy ids [B, T_c], alphas [T_c, n, B]."""
    B = video.shape[0]
    H = p['embed_att_Wa'].shape[0]
    dt = p['Wemb'].dtype
    emb, part = _frames(p, video.astype(dt))
    state = np.zeros((B, 2 * H), dt)
    h_prev = np.zeros((B, H), dt)
    current_embed = np.zeros((B, H), dt)
    words, saved_alphas, logits_all = [], [], []
    for i in range(n_caption_lstm_steps):
        alphas, atten = _attend(p, h_prev, emb, part)
        saved_alphas.append(alphas)
        out1, state = _lstm3(p, np.concatenate([atten, current_embed], 1), state)
        out2 = np.tanh(np.concatenate([out1, atten, current_embed], 1) @ p['embed_nn_Wp'] + p['embed_nn_bp'])
        h_prev = out1
        logit_words = out2 @ p['embed_word_W'] + p['embed_word_b']
        max_prob_index = logit_words.argmax(1)
        words.append(max_prob_index)
        logits_all.append(logit_words)
        current_embed = p['Wemb'][max_prob_index]
    ids = np.stack(words).T.astype(np.int64)
    if return_logits:
        return ids, np.stack(saved_alphas), np.stack(logits_all)
    return ids, np.stack(saved_alphas)
