"""TEST INFRASTRUCTURE — CPU oracle of one whole optimisation step (S1 + S2 + L1 + L2 + O1):
restates Model._train_on_batch (models/scene_flow_motion_field.py:152-227) functionally on state
dicts, with torch autograd on the CPU for the backward and an explicit Adam update
(torch.optim.Adam semantics, models/netinterface.py:96-97,127-129).

Also the `cpu_baseline` / `--impl reference` leg of bench.py (kind "port"): the Python reference cannot
travel to the GPU box, this restatement of the same PyTorch-CPU computation can.
"""
import torch

from . import depth_nets, geometry, sf_mlp


def _leafify(sd, trainable):
    out = {}
    for k, v in sd.items():
        if v.dtype.is_floating_point and trainable and not (k.endswith('running_mean') or k.endswith('running_var')):
            out[k] = v.detach().clone().requires_grad_()
        else:
            out[k] = v.detach()
    return out


def adam_update(p, g, state, lr, beta1=0.5, beta2=0.9, eps=1e-8):
    """One torch.optim.Adam step (amsgrad=False, weight_decay=0) for a single tensor; state is
    {'step','exp_avg','exp_avg_sq'} or {} on the first call."""
    if not state:
        state.update(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
    state['step'] += 1
    t = state['step']
    state['exp_avg'] = beta1 * state['exp_avg'] + (1 - beta1) * g
    state['exp_avg_sq'] = beta2 * state['exp_avg_sq'] + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
    denom = state['exp_avg_sq'].sqrt() / (bc2 ** 0.5) + eps
    return p - (lr / bc1) * state['exp_avg'] / denom


def train_step(sd_depth, sd_mlp, batch, opt, epoch, adam_depth=None, adam_mlp=None):
    """batch: tensors WITHOUT the DataLoader dim ([B,...]) on the CPU. Returns
    (batch_log, new_sd_depth, new_sd_mlp, extras). `opt` is the argparse namespace of the reference flags."""
    warm = epoch <= opt.warm_sf
    depth_t = _leafify(sd_depth, not warm)
    mlp_t = _leafify(sd_mlp, True)
    fwd = depth_nets.midas_forward if opt.midas else depth_nets.hourglass_forward
    ctx = torch.no_grad() if warm else torch.enable_grad()
    with ctx:
        d1 = fwd(depth_t, batch['img_1'])
        d2 = fwd(depth_t, batch['img_2'])
    R1, R2, t1, t2, K, Kinv = geometry._poses(batch)
    P1 = geometry.unproject(d1, R1, t1, Kinv)
    dt = float(batch['time_step'].reshape(-1)[0]) if torch.is_tensor(batch['time_step']) else float(batch['time_step'])
    gap = (batch['time_stamp_2'] - batch['time_stamp_1']).mean()
    steps = int((gap / dt).round().item())
    layers = sf_mlp.layers_from_state_dict(mlp_t)
    mkw = dict(n_freq_xyz=opt.n_freq_xyz, n_freq_t=opt.n_freq_t, time_dependent=opt.time_dependent,
               sf_mag_div=opt.sf_mag_div)
    ts1 = batch['time_stamp_1']
    sf = sf_mlp.sf_multi_step(P1, ts1, dt, steps, layers, **mkw)
    if opt.use_motion_seg:
        sf = sf * batch['motion_seg_1'].reshape(sf.shape[0], 1, *sf.shape[2:])
    loss, parts, r = geometry.reproject_and_loss(d1, d2, sf, batch, midas=opt.midas, warm=warm, use_disp=opt.use_disp,
                                                 use_disp_ratio=opt.use_disp_ratio, flow_mul=opt.flow_mul,
                                                 disp_mul=opt.disp_mul)
    total = loss * steps if opt.weight_steps else loss
    reg = None
    if opt.interp_steps > 0 and (not warm or opt.warm_reg) and opt.acc_mul > 0:
        reg = sf_mlp.acc_reg(P1, ts1, dt, layers, acc_mul=opt.acc_mul, **mkw)
        total = total + reg          # two backward() calls in the reference accumulate the same sum
    total.backward()
    log = {'size': opt.batch_size, 'loss': float(loss), 'total_loss': float(loss),
           'flow_loss_1_2': float(parts['flow_loss_1_2']), 'disp_loss_1_2': float(parts['disp_loss_1_2']),
           'sf_loss': float(parts['sf_loss']), 'acc_reg': float(reg) if reg is not None else 0}
    adam_depth = adam_depth if adam_depth is not None else {}
    adam_mlp = adam_mlp if adam_mlp is not None else {}
    new_depth, new_mlp = {}, {}
    for k, v in depth_t.items():
        if v.requires_grad and v.grad is not None:
            new_depth[k] = adam_update(v.detach(), v.grad, adam_depth.setdefault(k, {}), opt.lr, opt.adam_beta1, opt.adam_beta2)
        else:
            new_depth[k] = v.detach()
    for k, v in mlp_t.items():
        if v.requires_grad and v.grad is not None:
            new_mlp[k] = adam_update(v.detach(), v.grad, adam_mlp.setdefault(k, {}), opt.lr * opt.scene_lr_mul,
                                     opt.adam_beta1, opt.adam_beta2)
        else:
            new_mlp[k] = v.detach()
    extras = {'depth_1': d1.detach(), 'depth_2': d2.detach(), 'sf_1_2': sf.detach(), 'steps': steps,
              'grads_depth': {k: v.grad for k, v in depth_t.items() if v.requires_grad and v.grad is not None},
              'grads_mlp': {k: v.grad for k, v in mlp_t.items() if v.requires_grad and v.grad is not None}}
    return log, new_depth, new_mlp, extras
