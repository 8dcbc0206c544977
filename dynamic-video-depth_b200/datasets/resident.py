"""GPU-resident sequence cache + gap-bucketed, rank-disjoint pair sampler.

What it replaces in the reference: every training step `torch.load`s one pair file from disk
(datasets/davis_sequence.py:86-154), the DataLoader collates it, and `NetInterface.load_batch` copies 23 tensors to
the GPU (models/netinterface.py:163-177) - after the kernels are fast that is the step's critical path. A whole 80-200
frame sequence is small (images 1 MB/frame, flows + masks ~2 MB/pair at 384x224: < 1 GB), so it is loaded ONCE:

  ResidentSequence    pulls every pair of any reference-format Dataset (`__getitem__` -> the pair-file dict), de-duplicates
                      the frames, keeps images per FRAME and flows / masks / poses per PAIR on the device, and assembles the
                      batch dict the Model consumes by index_select - zero host->device bytes per step, no host sync (the
                      frame gap travels as `steps_hint`).
  GapBucketSampler    the DistributedSampler of train.py:301-305 made gap-aware: pairs are bucketed by frame gap (the number
                      of Euler sub-steps of the scene-flow chain = per-pair cost), every global step draws world*B pairs of
                      ONE gap, disjoint across ranks, reshuffled per epoch (`set_epoch`). Equal work per rank per step is what
                      keeps the gradient all-reduce from waiting on the slowest rank.
"""
import torch

# per-pair tensors of the pair-file format (leading dim = pairs in the file); everything else is rebuilt per batch
_PAIR_KEYS = ('flow_1_2', 'flow_2_1', 'mask_1', 'mask_2', 'motion_seg_1', 'R_1', 'R_1_T', 'R_2', 'R_2_T', 't_1', 't_2', 'K', 'K_inv',
              'depth_pred_1')


class ResidentSequence:
    def __init__(self, dataset, device, indices=None):
        """dataset[i] -> dict in the reference's pair-file format with ONE pair per item (leading dim 1 or absent)."""
        self.device = torch.device(device)
        idx = list(range(len(dataset))) if indices is None else list(indices)
        frames, self.pairs = {}, []
        per_pair = {k: [] for k in _PAIR_KEYS}
        self.time_step = None
        self.n_frames = None
        for i in idx:
            it = dataset[i]
            f1 = int(round(float(torch.as_tensor(it['frame_id_1']).reshape(-1)[0])))
            f2 = int(round(float(torch.as_tensor(it['frame_id_2']).reshape(-1)[0])))
            img1, img2 = self._one(it['img_1']), self._one(it['img_2'])
            frames.setdefault(f1, img1)
            frames.setdefault(f2, img2)
            ts = it['time_step']
            ts = float(ts.reshape(-1)[0]) if torch.is_tensor(ts) else float(ts)
            self.time_step = ts if self.time_step is None else self.time_step
            for k in _PAIR_KEYS:
                if k in it:
                    per_pair[k].append(self._one(torch.as_tensor(it[k])))
            t1 = float(torch.as_tensor(it['time_stamp_1']).reshape(-1)[0])
            self.pairs.append((f1, f2, t1, float(torch.as_tensor(it['time_stamp_2']).reshape(-1)[0])))
        ids = sorted(frames)
        self._slot = {f: j for j, f in enumerate(ids)}
        self.images = torch.stack([frames[f] for f in ids]).float().to(self.device)           # [F,3,H,W]
        self.H, self.W = self.images.shape[-2:]
        self.data = {k: torch.stack(v).float().to(self.device) for k, v in per_pair.items() if len(v) == len(self.pairs)}
        self._f1 = torch.tensor([self._slot[p[0]] for p in self.pairs], device=self.device)
        self._f2 = torch.tensor([self._slot[p[1]] for p in self.pairs], device=self.device)
        self._fid1 = torch.tensor([float(p[0]) for p in self.pairs], device=self.device)
        self._fid2 = torch.tensor([float(p[1]) for p in self.pairs], device=self.device)
        self._ts1 = torch.tensor([p[2] for p in self.pairs], device=self.device)
        self._ts2 = torch.tensor([p[3] for p in self.pairs], device=self.device)
        self.gaps = [int(round((p[3] - p[2]) / self.time_step)) for p in self.pairs]

    @staticmethod
    def _one(t):
        t = torch.as_tensor(t)
        return t[0] if (t.dim() > 0 and t.shape[0] == 1 and t.dim() in (4, 5, 6)) else t

    def __len__(self):
        return len(self.pairs)

    def nbytes(self):
        return self.images.numel() * 4 + sum(v.numel() * 4 for v in self.data.values())

    def batch(self, pair_indices):
        """Batch dict for `Model._train_on_batch` (no DataLoader dim: tensors are [B,...]), entirely on the device. All pairs
        must share their frame gap (GapBucketSampler guarantees it); it is passed on as `steps_hint`."""
        pi = list(pair_indices)
        gaps = {self.gaps[i] for i in pi}
        if len(gaps) != 1:
            raise ValueError('a resident batch must hold pairs of one frame gap (got %s)' % sorted(gaps))
        sel = torch.tensor(pi, device=self.device)
        B = len(pi)
        b = {k: v.index_select(0, sel) for k, v in self.data.items()}
        b['img_1'] = self.images.index_select(0, self._f1.index_select(0, sel))
        b['img_2'] = self.images.index_select(0, self._f2.index_select(0, sel))
        b['time_stamp_1'] = self._ts1.index_select(0, sel).view(B, 1, 1, 1).expand(B, 1, self.H, self.W).contiguous()
        b['time_stamp_2'] = self._ts2.index_select(0, sel).view(B, 1, 1, 1).expand(B, 1, self.H, self.W).contiguous()
        b['frame_id_1'] = self._fid1.index_select(0, sel)
        b['frame_id_2'] = self._fid2.index_select(0, sel)
        b['time_step'] = self.time_step                      # host float: no device read-back
        b['steps_hint'] = gaps.pop()
        b['pair_path'] = ['resident_%03d_%03d' % (self.pairs[i][0], self.pairs[i][1]) for i in pi]
        return b


class GapBucketSampler:
    """Iterates over lists of B pair indices for THIS rank. One global step = world*B pairs of one gap, taken from that gap's
    shuffled bucket; ranks receive disjoint slices; incomplete tails are dropped (drop_last of train.py:309-318)."""

    def __init__(self, gaps, pairs_per_step=1, world=1, rank=0, seed=0, shuffle=True):
        self.gaps = list(gaps)
        self.B, self.world, self.rank, self.seed, self.shuffle = int(pairs_per_step), int(world), int(rank), int(seed), shuffle
        self.epoch = 0
        self.buckets = {}
        for i, g in enumerate(self.gaps):
            self.buckets.setdefault(g, []).append(i)

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def _steps(self):
        g = torch.Generator().manual_seed(self.seed + 7919 * self.epoch)
        steps = []
        chunk = self.B * self.world
        for gap in sorted(self.buckets):
            idx = self.buckets[gap]
            if self.shuffle:
                idx = [idx[j] for j in torch.randperm(len(idx), generator=g).tolist()]
            for c in range(len(idx) // chunk):
                steps.append(idx[c * chunk:(c + 1) * chunk])
        if self.shuffle and steps:
            steps = [steps[j] for j in torch.randperm(len(steps), generator=g).tolist()]
        return steps

    def __len__(self):
        chunk = self.B * self.world
        return sum(len(v) // chunk for v in self.buckets.values())

    def __iter__(self):
        for glob in self._steps():
            yield glob[self.rank * self.B:(self.rank + 1) * self.B]


class ResidentLoader:
    """What `NetInterface.train_epoch` iterates over instead of a DataLoader: GPU-resident batches in sampler order."""

    def __init__(self, sequence, sampler):
        self.sequence, self.sampler = sequence, sampler

    def __len__(self):
        return len(self.sampler)

    def __iter__(self):
        for idx in self.sampler:
            yield self.sequence.batch(idx)
