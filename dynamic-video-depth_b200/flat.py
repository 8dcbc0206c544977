"""Flat fp32 parameter / gradient / Adam-moment buffers with per-tensor views under the reference's
state-dict names (SURVEY.md §8(b) checkpoint contract), the fused Adam step (O1) and the gradient
all-reduce of the data-parallel path (C1).

Reference behaviour being replaced:
  * two torch.optim.Adam instances looping over ~430 tensors (models/scene_flow_motion_field.py:113-115,212-213);
  * `DistributedDataParallel` wrappers that the reference builds and then DISCARDS (train.py:284-287), so
    its replicas never exchange gradients — here the intended mean all-reduce is done, on ONE flat buffer
    per net (NCCL over NVLink/NVSwitch), followed by the same Adam update on every rank;
  * one `dist.broadcast` per parameter tensor at start-up (train.py:290-292) → one flat broadcast.
"""
import ctypes

import torch

from . import _lib


def _align(n, a=64):
    return (n + a - 1) // a * a


class FlatParams:
    """Re-homes every parameter of `net` into one contiguous fp32 buffer (and its .grad into another).
    nn.Parameter objects are kept — `.data` / `.grad` become views — so `state_dict()` /
    `load_state_dict()` keep working with reference checkpoints."""

    def __init__(self, net, channels_last=False):
        """channels_last: conv weights with a spatial extent are stored O,kh,kw,I in the flat buffer (logical shape
        and state-dict values unchanged) so that cuDNN's NHWC kernels take them without a per-call conversion."""
        self.params = [p for p in net.parameters()]
        self.channels_last = channels_last
        dev = self.params[0].device
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += _align(p.numel(), 4)   # keep every tensor 16-byte aligned
        self.numel = _align(n, 4)
        self.data = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            view = self._view(self.data, p, o)
            view.copy_(p.data)
            p.data = view
            p.grad = self._view(self.grad, p, o)
            p._dvd_flat_grad = True     # kernels may accumulate straight into .grad (ops.BnAct, depth_engine): explicit opt-in

    def _view(self, buf, p, o):
        flat = buf[o:o + p.numel()]
        if self.channels_last and p.dim() == 4 and p.shape[2] * p.shape[3] > 1:
            O, I, kh, kw = p.shape
            return flat.view(O, kh, kw, I).permute(0, 3, 1, 2)
        return flat.view(p.shape)

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):   # autograd may have swapped the .grad tensor
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o or p.grad.stride() != p.data.stride():
                p.grad = self._view(self.grad, p, o)

    def broadcast(self, src=0):
        """One flat broadcast instead of one per tensor (train.py:290-292)."""
        import torch.distributed as dist
        dist.broadcast(self.data, src)

    def allreduce_grad(self, async_op=False):
        """Sum over ranks; the 1/world factor is folded into the Adam kernel (gscale)."""
        import torch.distributed as dist
        return dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, async_op=async_op)

    def bucket_ranges(self, net, starts):
        """Contiguous element ranges of the flat buffer delimited by the first parameter whose name starts with each prefix in
        `starts` (in parameter order): [0, s1), [s1, s2), ..., [s_last, numel). Used to all-reduce the gradient in buckets, in
        the order the explicit backward of depth_engine.MidasEngine completes them (decoder + layer4 first)."""
        names = [n for n, _ in net.named_parameters()]
        cuts = []
        for pre in starts:
            i = next(j for j, n in enumerate(names) if n.startswith(pre))
            cuts.append(self.offsets[i])
        cuts = [0] + cuts + [self.numel]
        return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]

    def allreduce_range(self, a, b, async_op=True):
        """Sum-all-reduce of grad[a:b]. With the NCCL backend the collective runs on the process group's own stream behind
        everything already enqueued on the current stream, i.e. it overlaps the backward kernels launched after this call;
        `work.wait()` later makes the current stream wait for it."""
        import torch.distributed as dist
        return dist.all_reduce(self.grad[a:b], op=dist.ReduceOp.SUM, async_op=async_op)


class FlatAdam:
    """torch.optim.Adam(betas, eps=1e-8, weight_decay=0, amsgrad=False) on a FlatParams buffer."""

    def __init__(self, flat, lr, betas=(0.9, 0.999), eps=1e-8):
        self.flat = flat
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        # the step counter also lives on the device ({int step, 1-b1^t, sqrt(1-b2^t), -}: dvd_adam_flat_dev increments it), so a
        # captured CUDA graph of the optimisation step replays with the right bias correction
        self.step_state = torch.zeros(4, dtype=torch.float32, device=flat.data.device)
        self.param_groups = [{'lr': self.lr, 'betas': self.betas, 'eps': self.eps, 'weight_decay': 0,
                              'amsgrad': False, 'params': list(range(len(flat.params)))}]

    def step(self, gscale=1.0):
        self.step_count += 1
        f = self.flat
        if f.data.is_cuda:
            lib = _lib.load()
            P = ctypes.c_void_p
            st = P(torch.cuda.current_stream().cuda_stream)
            from . import ops
            ops.LAUNCHES['n'] += 2
            _lib.check(lib.dvd_adam_flat_dev(P(f.data.data_ptr()), P(f.grad.data_ptr()), P(self.exp_avg.data_ptr()),
                                             P(self.exp_avg_sq.data_ptr()), f.numel, self.lr, self.betas[0], self.betas[1],
                                             self.eps, P(self.step_state.data_ptr()), float(gscale), st), 'dvd_adam_flat_dev')
        else:
            raise RuntimeError('FlatAdam needs CUDA buffers: dvd_b200 has no CPU compute path')

    def note_replayed(self, n=1):
        """A captured graph containing `step()` was replayed n times: keep the host-side counter (checkpoints) in step."""
        self.step_count += n

    def _sync_step_state(self):
        self.step_state.zero_()
        self.step_state[:1].view(torch.int32).fill_(int(self.step_count))

    # ---- torch.optim.Adam-compatible state dicts (models/netinterface.py:528-574) -----------------
    def state_dict(self):
        state = {}
        if self.step_count > 0:
            for i, (p, o) in enumerate(zip(self.flat.params, self.flat.offsets)):
                state[i] = {'step': torch.tensor(float(self.step_count)),
                            'exp_avg': self.flat._view(self.exp_avg, p, o).clone(memory_format=torch.contiguous_format),
                            'exp_avg_sq': self.flat._view(self.exp_avg_sq, p, o).clone(memory_format=torch.contiguous_format)}
        return {'state': state, 'param_groups': [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd, keep_training_params=True):
        steps = []
        for i, st in sd.get('state', {}).items():
            i = int(i)
            p, o = self.flat.params[i], self.flat.offsets[i]
            self.flat._view(self.exp_avg, p, o).copy_(st['exp_avg'])
            self.flat._view(self.exp_avg_sq, p, o).copy_(st['exp_avg_sq'])
            steps.append(int(float(st['step'])))
        self.step_count = max(steps) if steps else 0
        self._sync_step_state()
        if not keep_training_params and sd.get('param_groups'):
            g = sd['param_groups'][0]
            self.lr, self.betas, self.eps = float(g['lr']), tuple(g['betas']), float(g['eps'])
