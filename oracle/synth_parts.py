"""Restatement of the NON-decoder parts of the reference's synthesizer forward, for BASELINE.json configs[4]
("full synthesizer forward with synthetic HuBERT content features ... to measure the decoder's share").

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see the header of ``oracle/hifigan_oracle.py``): nothing under
``vcvits_b200/`` imports this file.  These modules are OUT of the accelerated hot path (only ``dec``
is in scope); they exist so that ``tools/full_synth.py`` can time a complete generator forward on the GPU box, where
``/root/reference`` is not available, with plain PyTorch standing in for the reference's own PyTorch code.

Each class keeps the reference's attribute names, so a reference ``state_dict`` loads unchanged
(``oracle/make_golden_synth.py`` pins every module against the reference class on the same weights):
  * ``LayerNorm``                     vits/model/modules.py:19-31
  * ``WN``                            vits/model/modules.py:109-183  (gated dilated conv stack;
                                      ``fused_add_tanh_sigmoid_multiply`` vits/commons.py:99-106)
  * ``ResidualCouplingLayer``/``Flip`` vits/model/modules.py:263-273,285-333
  * ``ResidualCouplingBlock``         vits/model/flow.py:7-38
  * ``PosteriorEncoder``              vits/model/encoders/posterior_encoder.py:9-39
  * ``MultiHeadAttention``/``FFN``/``TransformerEncoder``  vits/model/transformer/relative_attention_transformer.py:13-47,103-311
  * ``PreloadHubertContentEncoder``   vits/model/encoders/content_encoder.py:76-126
  * ``sequence_mask``/``slice_segments``/``rand_slice_segments``  vits/commons.py:48-64,126-130
  * ``generator_forward``             vits/model/synthesizers/synthesizer_svc.py:70-88 (SynthesizerSVC.forward)
"""
from __future__ import annotations

import math
import warnings

import torch
from torch import nn
from torch.nn import functional as F


def _wn(m):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(m, name="weight")


def sequence_mask(length, max_length=None):
    if max_length is None:
        max_length = length.max()
    x = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def slice_segments(x, ids_str, segment_size=4):
    ret = torch.zeros_like(x[:, :, :segment_size])
    for i in range(x.size(0)):
        idx_str = ids_str[i]
        ret[i] = x[i, :, idx_str:idx_str + segment_size]
    return ret


def rand_slice_segments(x, x_lengths=None, segment_size=4):
    b, d, t = x.size()
    if x_lengths is None:
        x_lengths = t
    ids_str_max = x_lengths - segment_size + 1
    ids_str = (torch.rand([b], device=x_lengths.device) * ids_str_max).to(dtype=torch.long)
    return slice_segments(x, ids_str, segment_size), ids_str


class LayerNorm(nn.Module):
    def __init__(self, channels, eps=1e-5):
        super().__init__()
        self.channels, self.eps = channels, eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))

    def forward(self, x):
        x = x.transpose(1, -1)
        x = F.layer_norm(x, (self.channels,), self.gamma, self.beta, self.eps)
        return x.transpose(1, -1)


class WN(nn.Module):
    def __init__(self, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0, p_dropout=0):
        super().__init__()
        self.hidden_channels, self.n_layers, self.gin_channels = hidden_channels, n_layers, gin_channels
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        self.drop = nn.Dropout(p_dropout)
        if gin_channels != 0:
            self.cond_layer = _wn(nn.Conv1d(gin_channels, 2 * hidden_channels * n_layers, 1))
        for i in range(n_layers):
            dilation = dilation_rate ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(_wn(nn.Conv1d(hidden_channels, 2 * hidden_channels, kernel_size, dilation=dilation,
                                                padding=padding)))
            res_skip_channels = 2 * hidden_channels if i < n_layers - 1 else hidden_channels
            self.res_skip_layers.append(_wn(nn.Conv1d(hidden_channels, res_skip_channels, 1)))

    def forward(self, x, x_mask, g=None):
        output = torch.zeros_like(x)
        h = self.hidden_channels
        if g is not None:
            g = self.cond_layer(g)
        for i in range(self.n_layers):
            x_in = self.in_layers[i](x)
            g_l = g[:, i * 2 * h:(i + 1) * 2 * h, :] if g is not None else torch.zeros_like(x_in)
            in_act = x_in + g_l
            acts = self.drop(torch.tanh(in_act[:, :h, :]) * torch.sigmoid(in_act[:, h:, :]))
            res_skip_acts = self.res_skip_layers[i](acts)
            if i < self.n_layers - 1:
                x = (x + res_skip_acts[:, :h, :]) * x_mask
                output = output + res_skip_acts[:, h:, :]
            else:
                output = output + res_skip_acts
        return output * x_mask


class Flip(nn.Module):
    def forward(self, x, *args, reverse=False, **kwargs):
        x = torch.flip(x, [1])
        if not reverse:
            return x, torch.zeros(x.size(0)).to(dtype=x.dtype, device=x.device)
        return x


class ResidualCouplingLayer(nn.Module):
    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, n_layers, p_dropout=0, gin_channels=0,
                 mean_only=False):
        super().__init__()
        self.half_channels = channels // 2
        self.mean_only = mean_only
        self.pre = nn.Conv1d(self.half_channels, hidden_channels, 1)
        self.enc = WN(hidden_channels, kernel_size, dilation_rate, n_layers, p_dropout=p_dropout, gin_channels=gin_channels)
        self.post = nn.Conv1d(hidden_channels, self.half_channels * (2 - mean_only), 1)
        self.post.weight.data.zero_()
        self.post.bias.data.zero_()

    def forward(self, x, x_mask, g=None, reverse=False):
        x0, x1 = torch.split(x, [self.half_channels] * 2, 1)
        h = self.pre(x0) * x_mask
        h = self.enc(h, x_mask, g=g)
        stats = self.post(h) * x_mask
        if not self.mean_only:
            m, logs = torch.split(stats, [self.half_channels] * 2, 1)
        else:
            m, logs = stats, torch.zeros_like(stats)
        if not reverse:
            x1 = m + x1 * torch.exp(logs) * x_mask
            return torch.cat([x0, x1], 1), torch.sum(logs, [1, 2])
        x1 = (x1 - m) * torch.exp(-logs) * x_mask
        return torch.cat([x0, x1], 1)


class ResidualCouplingBlock(nn.Module):
    def __init__(self, channels, hidden_channels, kernel_size, dilation_rate, n_layers, n_flows=4, gin_channels=0):
        super().__init__()
        self.flows = nn.ModuleList()
        for _ in range(n_flows):
            self.flows.append(ResidualCouplingLayer(channels, hidden_channels, kernel_size, dilation_rate, n_layers,
                                                    gin_channels=gin_channels, mean_only=True))
            self.flows.append(Flip())

    def forward(self, x, x_mask, g=None, reverse=False):
        if not reverse:
            for flow in self.flows:
                x, _ = flow(x, x_mask, g=g, reverse=reverse)
        else:
            for flow in reversed(self.flows):
                x = flow(x, x_mask, g=g, reverse=reverse)
        return x


class PosteriorEncoder(nn.Module):
    def __init__(self, in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=0):
        super().__init__()
        self.out_channels = out_channels
        self.pre = nn.Conv1d(in_channels, hidden_channels, 1)
        self.enc = WN(hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels=gin_channels)
        self.proj = nn.Conv1d(hidden_channels, out_channels * 2, 1)

    def forward(self, x, x_lengths, g=None):
        x_mask = torch.unsqueeze(sequence_mask(x_lengths, x.size(2)), 1).to(x.dtype)
        x = self.pre(x) * x_mask
        x = self.enc(x, x_mask, g=g)
        stats = self.proj(x) * x_mask
        m, logs = torch.split(stats, self.out_channels, dim=1)
        z = (m + torch.randn_like(m) * torch.exp(logs)) * x_mask
        return z, m, logs, x_mask


class MultiHeadAttention(nn.Module):
    """Windowed relative-position self-attention (window_size given, heads share the relative embeddings)."""

    def __init__(self, channels, out_channels, n_heads, p_dropout=0., window_size=None):
        super().__init__()
        assert channels % n_heads == 0
        self.n_heads, self.window_size = n_heads, window_size
        self.k_channels = channels // n_heads
        self.conv_q = nn.Conv1d(channels, channels, 1)
        self.conv_k = nn.Conv1d(channels, channels, 1)
        self.conv_v = nn.Conv1d(channels, channels, 1)
        self.conv_o = nn.Conv1d(channels, out_channels, 1)
        self.drop = nn.Dropout(p_dropout)
        if window_size is not None:
            rel_stddev = self.k_channels ** -0.5
            self.emb_rel_k = nn.Parameter(torch.randn(1, window_size * 2 + 1, self.k_channels) * rel_stddev)
            self.emb_rel_v = nn.Parameter(torch.randn(1, window_size * 2 + 1, self.k_channels) * rel_stddev)
        nn.init.xavier_uniform_(self.conv_q.weight)
        nn.init.xavier_uniform_(self.conv_k.weight)
        nn.init.xavier_uniform_(self.conv_v.weight)

    def forward(self, x, c, attn_mask=None):
        q, k, v = self.conv_q(x), self.conv_k(c), self.conv_v(c)
        b, d, t_s, t_t = (*k.size(), q.size(2))
        q = q.view(b, self.n_heads, self.k_channels, t_t).transpose(2, 3)
        k = k.view(b, self.n_heads, self.k_channels, t_s).transpose(2, 3)
        v = v.view(b, self.n_heads, self.k_channels, t_s).transpose(2, 3)
        scores = torch.matmul(q / math.sqrt(self.k_channels), k.transpose(-2, -1))
        if self.window_size is not None:
            rel_k = self._rel_emb(self.emb_rel_k, t_s)
            rel_logits = torch.matmul(q / math.sqrt(self.k_channels), rel_k.unsqueeze(0).transpose(-2, -1))
            scores = scores + self._rel_to_abs(rel_logits)
        if attn_mask is not None:
            scores = scores.masked_fill(attn_mask == 0, -1e4)
        p_attn = self.drop(F.softmax(scores, dim=-1))
        out = torch.matmul(p_attn, v)
        if self.window_size is not None:
            rel_w = self._abs_to_rel(p_attn)
            out = out + torch.matmul(rel_w, self._rel_emb(self.emb_rel_v, t_s).unsqueeze(0))
        out = out.transpose(2, 3).contiguous().view(b, d, t_t)
        return self.conv_o(out)

    def _rel_emb(self, emb, length):
        pad_length = max(length - (self.window_size + 1), 0)
        start = max((self.window_size + 1) - length, 0)
        if pad_length > 0:
            emb = F.pad(emb, [0, 0, pad_length, pad_length, 0, 0])
        return emb[:, start:start + 2 * length - 1]

    @staticmethod
    def _rel_to_abs(x):
        b, h, l, _ = x.size()
        x = F.pad(x, [0, 1, 0, 0, 0, 0, 0, 0])
        x_flat = F.pad(x.view([b, h, l * 2 * l]), [0, l - 1, 0, 0, 0, 0])
        return x_flat.view([b, h, l + 1, 2 * l - 1])[:, :, :l, l - 1:]

    @staticmethod
    def _abs_to_rel(x):
        b, h, l, _ = x.size()
        x = F.pad(x, [0, l - 1, 0, 0, 0, 0, 0, 0])
        x_flat = F.pad(x.view([b, h, l ** 2 + l * (l - 1)]), [l, 0, 0, 0, 0, 0])
        return x_flat.view([b, h, l, 2 * l])[:, :, :, 1:]


class FFN(nn.Module):
    def __init__(self, in_channels, out_channels, filter_channels, kernel_size, p_dropout=0.):
        super().__init__()
        self.kernel_size = kernel_size
        self.conv_1 = nn.Conv1d(in_channels, filter_channels, kernel_size)
        self.conv_2 = nn.Conv1d(filter_channels, out_channels, kernel_size)
        self.drop = nn.Dropout(p_dropout)

    def _pad(self, x):
        if self.kernel_size == 1:
            return x
        return F.pad(x, [(self.kernel_size - 1) // 2, self.kernel_size // 2, 0, 0, 0, 0])

    def forward(self, x, x_mask):
        x = self.drop(torch.relu(self.conv_1(self._pad(x * x_mask))))
        return self.conv_2(self._pad(x * x_mask)) * x_mask


class TransformerEncoder(nn.Module):
    def __init__(self, hidden_channels, filter_channels, n_heads, n_layers, kernel_size=1, p_dropout=0., window_size=4):
        super().__init__()
        self.n_layers = n_layers
        self.drop = nn.Dropout(p_dropout)
        self.attn_layers, self.norm_layers_1 = nn.ModuleList(), nn.ModuleList()
        self.ffn_layers, self.norm_layers_2 = nn.ModuleList(), nn.ModuleList()
        for _ in range(n_layers):
            self.attn_layers.append(MultiHeadAttention(hidden_channels, hidden_channels, n_heads, p_dropout=p_dropout,
                                                       window_size=window_size))
            self.norm_layers_1.append(LayerNorm(hidden_channels))
            self.ffn_layers.append(FFN(hidden_channels, hidden_channels, filter_channels, kernel_size, p_dropout=p_dropout))
            self.norm_layers_2.append(LayerNorm(hidden_channels))

    def forward(self, x, x_mask):
        attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
        x = x * x_mask
        for i in range(self.n_layers):
            y = self.drop(self.attn_layers[i](x, x, attn_mask))
            x = self.norm_layers_1[i](x + y)
            y = self.drop(self.ffn_layers[i](x, x_mask))
            x = self.norm_layers_2[i](x + y)
        return x * x_mask


class PreloadHubertContentEncoder(nn.Module):
    def __init__(self, out_channels, hidden_channels, filter_channels, n_heads, n_layers, kernel_size, p_dropout,
                 hubert_channels, num_pitch):
        super().__init__()
        self.out_channels = out_channels
        proj_channels = hidden_channels // 2
        self.hubert_proj = nn.Linear(hubert_channels, proj_channels)
        self.emb_pitch = nn.Embedding(num_pitch, proj_channels)
        nn.init.normal_(self.emb_pitch.weight, 0.0, proj_channels ** -0.5)
        self.pitch_proj = nn.Linear(proj_channels, proj_channels)
        self.encoder = TransformerEncoder(hidden_channels, filter_channels, n_heads, n_layers, kernel_size, p_dropout)
        self.proj = nn.Conv1d(hidden_channels, out_channels * 2, 1)

    def forward(self, x, x_lengths, pitch, pitch_lengths):
        hubert_out = self.hubert_proj(x.transpose(1, -1)).transpose(1, -1)
        pitch_out = self.pitch_proj(self.emb_pitch(pitch)).transpose(1, -1)
        out = torch.concat((hubert_out, pitch_out), dim=1)
        x_mask = torch.unsqueeze(sequence_mask(x_lengths.int(), out.size(2)), 1).to(x.dtype)
        x_out = self.encoder(out * x_mask, x_mask)
        stats = self.proj(x_out) * x_mask
        m, logs = torch.split(stats, self.out_channels, dim=1)
        return x_out, m, logs, x_mask


# configs/base.json:44-68 (model section) + data section values used by vits/light/vcvits.py:33-37
BASE_SYNTH = dict(spec_channels=1025, segment_frames=32, inter_channels=256, hidden_channels=256, filter_channels=768,
                  n_heads=4, n_layers=3, kernel_size=3, p_dropout=0.1, hubert_channels=1280, num_pitch=512,
                  n_speakers=512, gin_channels=256)


class GeneratorParts(nn.Module):
    """enc_p / enc_q / flow / emb_g of SynthesizerSVC (synthesizer_svc.py:56-68) -- everything but ``dec``."""

    def __init__(self, c: dict):
        super().__init__()
        self.c = c
        self.enc_p = PreloadHubertContentEncoder(c["inter_channels"], c["hidden_channels"], c["filter_channels"], c["n_heads"],
                                                 c["n_layers"], c["kernel_size"], c["p_dropout"], c["hubert_channels"],
                                                 c["num_pitch"])
        self.enc_q = PosteriorEncoder(c["spec_channels"], c["inter_channels"], c["hidden_channels"], 5, 1, 16,
                                      gin_channels=c["gin_channels"])
        self.flow = ResidualCouplingBlock(c["inter_channels"], c["hidden_channels"], 5, 1, 4, gin_channels=c["gin_channels"])
        self.emb_g = nn.Embedding(c["n_speakers"], c["gin_channels"])


def generator_forward(parts: GeneratorParts, dec, feats, feat_lengths, pitch, spec, spec_lengths, sid, timer=None):
    """SynthesizerSVC.forward (synthesizer_svc.py:70-88) with pre-extracted content features; ``dec`` is any decoder
    module with the Generator call contract.  ``timer(name)`` is called before each part (None = no timing)."""
    tick = timer if timer is not None else (lambda name: None)
    tick("enc_p")
    x, m_p, logs_p, x_mask = parts.enc_p(feats, feat_lengths, pitch, feat_lengths)
    g = parts.emb_g(sid).unsqueeze(-1)
    tick("enc_q")
    z, m_q, logs_q, y_mask = parts.enc_q(spec, spec_lengths, g=g)
    tick("flow")
    z_p = parts.flow(z, y_mask, g=g)
    tick("interp+slice")
    m_p = F.interpolate(m_p, size=(spec.shape[2],), mode="nearest")
    logs_p = F.interpolate(logs_p, size=(spec.shape[2],), mode="nearest")
    z_slice, ids_slice = rand_slice_segments(z, spec_lengths, parts.c["segment_frames"])
    tick("dec")
    o = dec(z_slice, g=g) if getattr(dec, "gin_channels", 0) else dec(z_slice)
    tick(None)
    return o, ids_slice, z_slice, x_mask, y_mask, (z, z_p, m_p, logs_p, m_q, logs_q)
