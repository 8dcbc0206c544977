"""TEST INFRASTRUCTURE — CPU oracle of the scene-flow MLP (M1–M4) and the acceleration regulariser (L2).

Restates, with plain torch tensor ops on the CPU (any float dtype):
  periodic_embed       networks/blocks.py:19-34       (PeriodicEmbed.forward)
  mlp_forward          networks/sceneflow_field.py:20-53 + networks/blocks.py:50-102 (1×1 convs + LeakyReLU 0.2)
  sf_net               models/scene_flow_motion_field.py:346-358 (forward_sf_net: ÷ sf_mag_div)
  sf_multi_step        models/scene_flow_motion_field.py:360-367 (Euler advection)
  acc_reg              models/scene_flow_motion_field.py:326-344 (_opt_reg)
Weights are passed as a list [(W_l [out,in], b_l [out])] in layer order (state-dict names
convs.{l}.conv.{weight,bias}, weight stored [out,in,1,1]).
"""
import torch


def freqs(n_freq, dtype=torch.float32):
    # linspace(1, max_freq+1, steps=N_freq) with max_freq = N_freq (networks/blocks.py:23-24;
    # ctor args networks/sceneflow_field.py:29,33). Always generated in fp32 like the reference,
    # then cast, so the fp64 oracle uses the same frequencies.
    return torch.linspace(1, n_freq + 1, steps=n_freq, dtype=torch.float32).to(dtype)


def periodic_embed(x, n_freq):
    """x [B,C,H,W] → [B, C·(1+2·n_freq), H, W]: [x, cos(f_k x) for k, sin(f_k x) for k]."""
    f = freqs(n_freq, x.dtype).to(x.device)
    out = [x]
    for fn in (torch.cos, torch.sin):
        for k in range(n_freq):
            out.append(fn(f[k] * x))
    return torch.cat(out, 1)


def mlp_forward(xyz, t, layers, n_freq_xyz=16, n_freq_t=16, time_dependent=True, probe=None):
    """Raw network output [B,3,H,W] (before ÷ sf_mag_div). `probe` (list) collects, per hidden layer,
    min_c |pre-activation| per pixel — used by the tests to find pixels sitting on a LeakyReLU kink."""
    feat = periodic_embed(xyz, n_freq_xyz) if n_freq_xyz > 0 else xyz
    if time_dependent:
        te = periodic_embed(t, n_freq_t) if n_freq_t > 0 else t
        feat = torch.cat([te, feat], 1)  # t first (networks/sceneflow_field.py:50)
    h = feat
    for i, (w, b) in enumerate(layers):
        h = torch.einsum('oi,bihw->bohw', w.reshape(w.shape[0], -1), h) + b.reshape(1, -1, 1, 1)
        if i + 1 < len(layers):
            if probe is not None:
                probe.append(h.detach().abs().amin(dim=1))
            h = torch.nn.functional.leaky_relu(h, 0.2)
    return h


def sf_net(xyz, t, layers, sf_mag_div=100.0, **kw):
    return mlp_forward(xyz, t, layers, **kw) / sf_mag_div


def sf_multi_step(p, t, time_step, steps, layers, **kw):
    """sf_acc = Σ_i s_i with p ← p + s_i, t ← t + Δt."""
    acc = torch.zeros_like(p)
    for _ in range(steps):
        s = sf_net(p, t, layers, **kw)
        acc = acc + s
        p = p + s
        t = t + time_step
    return acc


def acc_reg(p, t, time_step, layers, acc_mul=1.0, **kw):
    """acc_mul · Σ|s1 − s0| / (numel + 1e-6)."""
    s0 = sf_net(p, t, layers, **kw)
    s1 = sf_net(p + s0, t + time_step, layers, **kw)
    return acc_mul * (s1 - s0).abs().sum() / (s0.numel() + 1e-6)


def layers_from_state_dict(sd, prefix='', dtype=None):
    out = []
    i = 0
    while f'{prefix}convs.{i}.conv.weight' in sd:
        w = sd[f'{prefix}convs.{i}.conv.weight']
        b = sd[f'{prefix}convs.{i}.conv.bias']
        if dtype is not None:
            w, b = w.to(dtype), b.to(dtype)
        out.append((w.reshape(w.shape[0], -1), b))
        i += 1
    return out


def kink_band(p, t, time_step, n_eval, layers, width=3e-5, **kw):
    """Pixels [B,H,W] (bool) where some hidden pre-activation of some Euler step lies within
    +-width of the LeakyReLU kink. There the derivative is discontinuous, so two correct fp32
    implementations may legitimately pick different slopes (the reference on CPU vs on GPU does too);
    the parity tests zero the cotangent on these (independent, 1x1-conv) pixels."""
    band = None
    p, t = p.double(), t.double()
    l64 = [(w.double(), b.double()) for w, b in layers]
    for _ in range(n_eval):
        probe = []
        s = mlp_forward(p, t, l64, probe=probe, **{k: v for k, v in kw.items() if k != 'sf_mag_div'}) / kw.get('sf_mag_div', 100.0)
        m = torch.stack(probe, 0).amin(0) < width
        band = m if band is None else (band | m)
        p = p + s
        t = t + time_step
    return band


def init_layers(n_in=132, width=256, n_hidden=4, n_out=3, seed=0, dtype=torch.float32):
    """kaiming_normal_(a=0.2, fan_in), bias 0 (models/scene_flow_motion_field.py:123;
    models/netinterface.py:55-84). Seeded stand-alone variant used by tests/bench."""
    g = torch.Generator().manual_seed(seed)
    dims = [n_in] + [width] * (n_hidden + 1) + [n_out]
    layers = []
    for i in range(len(dims) - 1):
        gain = (2.0 / (1 + 0.2 ** 2)) ** 0.5
        std = gain / dims[i] ** 0.5
        w = torch.randn(dims[i + 1], dims[i], generator=g, dtype=torch.float64) * std
        layers.append((w.to(dtype), torch.zeros(dims[i + 1], dtype=dtype)))
    return layers
