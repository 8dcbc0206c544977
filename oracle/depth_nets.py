"""TEST INFRASTRUCTURE — CPU oracle of the depth nets (D1 / D1' / D2), state-dict driven and purely
functional (no nn.Module): an independent restatement of

  MidasNet.forward                  third_party/MiDaS.py:206-246, third_party/midas_blocks.py:28-168
  ResNeXt101-32x8d encoder          torch.hub facebookresearch/WSL-Images (midas_blocks.py:48-50) ==
                                    torchvision ResNet(Bottleneck,[3,4,23,3],groups=32,width_per_group=8)
  HourglassModel_Embed.forward      third_party/hourglass.py:21-212

`sd` maps the reference's state-dict names to tensors (leaf tensors may require grad: the oracle's
backward is torch autograd on the CPU). BatchNorm is ALWAYS in eval mode
(models/scene_flow_motion_field.py:157,168; hourglass.py:200-208). Pinned against the reference by
tests/test_oracle_vs_reference.py and tests/golden/step_golden.pt.
"""
import torch
import torch.nn.functional as F


def _bn(sd, pre, x, affine=True):
    return F.batch_norm(x, sd[pre + '.running_mean'], sd[pre + '.running_var'],
                        sd[pre + '.weight'] if affine else None, sd[pre + '.bias'] if affine else None,
                        training=False, eps=1e-5)


def _conv(sd, pre, x, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[pre + '.weight'], sd.get(pre + '.bias'), stride=stride, padding=padding, groups=groups)


# ---- ResNeXt101-32x8d -----------------------------------------------------------------------------------
def _bottleneck(sd, pre, x, stride):
    idt = x
    if (pre + '.downsample.0.weight') in sd:
        idt = _bn(sd, pre + '.downsample.1', _conv(sd, pre + '.downsample.0', x, stride=stride))
    y = F.relu(_bn(sd, pre + '.bn1', _conv(sd, pre + '.conv1', x)))
    y = F.relu(_bn(sd, pre + '.bn2', _conv(sd, pre + '.conv2', y, stride=stride, padding=1, groups=32)))
    y = _bn(sd, pre + '.bn3', _conv(sd, pre + '.conv3', y))
    return F.relu(y + idt)


def _stage(sd, pre, x, blocks, stride):
    for i in range(blocks):
        x = _bottleneck(sd, '%s.%d' % (pre, i), x, stride if i == 0 else 1)
    return x


def _rcu(sd, pre, x):
    r = F.relu(x)   # in-place ReLU in the reference: the skip connection sees relu(x) (midas_blocks.py:121,130-135)
    y = _conv(sd, pre + '.conv1', r, padding=1)
    y = _conv(sd, pre + '.conv2', F.relu(y), padding=1)
    return y + r


def _fusion(sd, pre, a, b=None):
    out = a if b is None else a + _rcu(sd, pre + '.resConfUnit1', b)
    out = _rcu(sd, pre + '.resConfUnit2', out)
    return F.interpolate(out, scale_factor=2, mode='bilinear', align_corners=True)


def midas_forward(sd, x, normalize_input=True, resize=None):
    if normalize_input:
        mean = torch.tensor([0.485, 0.456, 0.406], dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
        x = (x - mean) / std
    orig = x.shape[-2:]
    if resize is not None:
        x = F.interpolate(x, size=resize, mode='bicubic', align_corners=True)
    p = 'pretrained.'
    y = F.relu(_bn(sd, p + 'layer1.1', _conv(sd, p + 'layer1.0', x, stride=2, padding=3)))
    y = F.max_pool2d(y, 3, stride=2, padding=1)
    l1 = _stage(sd, p + 'layer1.4', y, 3, 1)
    l2 = _stage(sd, p + 'layer2', l1, 4, 2)
    l3 = _stage(sd, p + 'layer3', l2, 23, 2)
    l4 = _stage(sd, p + 'layer4', l3, 3, 2)
    s = 'scratch.'
    r = [_conv(sd, s + 'layer%d_rn' % (i + 1), l, padding=1) for i, l in enumerate((l1, l2, l3, l4))]
    path = _fusion(sd, s + 'refinenet4', r[3])
    path = _fusion(sd, s + 'refinenet3', path, r[2])
    path = _fusion(sd, s + 'refinenet2', path, r[1])
    path = _fusion(sd, s + 'refinenet1', path, r[0])
    o = _conv(sd, s + 'output_conv.0', path, padding=1)
    o = F.interpolate(o, scale_factor=2, mode='bilinear', align_corners=False)
    o = F.relu(_conv(sd, s + 'output_conv.2', o, padding=1))
    o = F.relu(_conv(sd, s + 'output_conv.4', o))
    o = 10000.0 / torch.clamp(o, min=1e-2)
    if resize is not None:
        o = F.interpolate(o, size=orig, mode='bicubic', align_corners=True)
    return o


# ---- hourglass --------------------------------------------------------------------------------------------
_SPECS = {
    'E': [[64], [3, 32, 64], [5, 32, 64], [7, 32, 64]], 'F': [[64], [3, 64, 64], [7, 64, 64], [11, 64, 64]],
    'B': [[32], [3, 32, 32], [5, 32, 32], [7, 32, 32]], 'C': [[32], [3, 64, 32], [7, 64, 32], [11, 64, 32]],
    'A': [[16], [3, 64, 16], [7, 64, 16], [11, 64, 16]], 'G': [[32], [3, 32, 32], [5, 32, 32], [7, 32, 32]],
    'H': [[32], [3, 64, 32], [5, 64, 32], [7, 64, 32]], 'I': [[16], [3, 32, 16], [7, 32, 16], [11, 32, 16]],
}
# each level = (branch0, branch1); entries: 'P' avg-pool, 'U' bilinear x2 (align_corners=True), 'n' nested level n,
# letters = inception specs (third_party/hourglass.py:60-158)
_LEVELS = {
    1: (['E', 'E'], ['P', 'E', 'E', 'E', 'U']),
    2: (['E', 'F'], ['P', 'E', 'E', 1, 'E', 'F', 'U']),
    3: (['P', 'B', 'E', 2, 'E', 'G', 'U'], ['B', 'C']),
    4: (['P', 'B', 'B', 3, 'H', 'I', 'U'], ['A']),
}


def _inception(sd, pre, x, spec):
    outs = [F.relu(_bn(sd, pre + '.convs.0.1', _conv(sd, pre + '.convs.0.0', x), affine=False))]
    for i, (k, _mid, _out) in enumerate(spec[1:], 1):
        b = pre + '.convs.%d' % i
        y = F.relu(_bn(sd, b + '.1', _conv(sd, b + '.0', x), affine=False))
        y = F.relu(_bn(sd, b + '.4', _conv(sd, b + '.3', y, padding=(k - 1) // 2), affine=False))
        outs.append(y)
    return torch.cat(outs, 1)


def _level(sd, pre, x, n):
    res = []
    for bi, ops in enumerate(_LEVELS[n]):
        y = x
        for oi, op in enumerate(ops):
            name = '%s.list.%d.%d' % (pre, bi, oi)
            if op == 'P':
                y = F.avg_pool2d(y, 2)
            elif op == 'U':
                y = F.interpolate(y, scale_factor=2, mode='bilinear', align_corners=True)
            elif isinstance(op, int):
                y = _level(sd, name, y, op)
            else:
                y = _inception(sd, name, y, _SPECS[op])
        res.append(y)
    return res[0] + res[1]


def hourglass_forward(sd, x, noexp=False):
    p = 'net_depth.'
    y = F.relu(_bn(sd, p + 'seq.1', _conv(sd, p + 'seq.0', x, padding=3)))
    y = _level(sd, p + 'seq.3', y, 4)
    pred = _conv(sd, p + 'pred_layer', y, padding=1)
    return pred if noexp else torch.exp(pred)
