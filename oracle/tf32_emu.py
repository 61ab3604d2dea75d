"""ORACLE (test infrastructure, not product code) -- the FOA offset head with the tensor cores'
TF32 operand rounding emulated on the CPU, forward AND backward.

Why it exists: the GPU's gradients of the 10-deep FOA conv stacks deviate
3-13 % from the fp32 oracle.  The claim is that this is inherent to TF32 operands -- every layer
output is rounded to a 10-bit mantissa before the next tensor-core op reads it, the rounding noise
(2.4e-4 relative per value) flips the ReLU mask of activations that are within ~1e-3 of zero, and
ten layers of that move a weight gradient by several percent -- and not a defect of the fused
backward.  This module restates exactly what the product computes for
``OffsetHeadExpandFeature`` (mmdet/models/roi_heads/attribute_heads/
offset_head_expand_feature.py:134-161 forward, :198-205 loss), with the product's rounding points:

  forward   a_0 = rot90(x, k)                                   (exact permutation)
            a_l = rna(relu(conv3x3(a_{l-1}, rna(W_l)) + b_l))   l = 1..10, per branch
            h_i = rna(relu(linear(h_{i-1}, rna(V_i)) + c_i))    i = 1, 2 (shared by the branches)
            o   = linear(h_2, rna(U)) + d                        (not rounded: feeds the loss)
  backward  g_o = dL/do (fp32; the tensor core TRUNCATES this one unrounded operand)
            dW  = g^T a_in, db = sum g, g_in = rna((a_in > 0) * (g W^r))  layer by layer

fp32 accumulation everywhere (as the tensor core does; only the ORDER of the additions differs,
~1e-6 relative).  `rna` = round-to-nearest, ties away from zero, to 10 mantissa bits
(cvt.rna.tf32.f32).  With ``emulate=False`` the same code is the plain fp32 head, so the two can
be compared on identical inputs without any GPU: tests/test_oracle_tf32.py.
"""
import torch
import torch.nn.functional as F


def rna(t):
    """cvt.rna.tf32.f32: round to nearest (ties away) keeping 10 mantissa bits."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def trunc(t):
    """What tcgen05.mma kind::tf32 does to an operand that was not pre-rounded."""
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def _ident(t):
    return t


def smooth_l1_grad(pred, target, beta, scale):
    d = pred - target
    return torch.where(d.abs() < beta, d / beta, torch.sign(d)) * scale


def foa_head_forward_backward(x, p, targets, prefix='roi_head.offset_head', rotations=(0, 90, 180, 270),
                              num_convs=10, loss_weight=16.0, beta=1.0, emulate=True):
    """x [P,256,7,7] RoI features (what the offset RoI extractor produced), p: state_dict,
    targets [4P,2] branch-major.  Returns (loss, {param name: gradient}) of
    loss_weight * mean(smooth_l1(offset_pred - targets)) over all 8P elements."""
    r = rna if emulate else _ident
    t = trunc if emulate else _ident
    P = x.shape[0]
    nb = len(rotations)
    grads = {}
    # ---------------------------------------------------------------- forward
    acts = []                      # per branch: [a_0 .. a_10]
    for bi, rot in enumerate(rotations):
        a = torch.rot90(x, rot // 90, dims=(2, 3))
        seq = [r(a)]
        for c in range(num_convs):
            w = r(p[f'{prefix}.expand_convs.{bi}.{c}.weight'])
            b = p[f'{prefix}.expand_convs.{bi}.{c}.bias']
            seq.append(r(F.relu(F.conv2d(seq[-1], w, b, padding=1))))
        acts.append(seq)
    h0 = torch.cat([s[-1].reshape(P, -1) for s in acts], 0)          # [4P, 12544], branch-major
    V = [r(p[f'{prefix}.fcs.{i}.weight']) for i in range(2)]
    c = [p[f'{prefix}.fcs.{i}.bias'] for i in range(2)]
    h1 = r(F.relu(F.linear(h0, V[0], c[0])))
    h2 = r(F.relu(F.linear(h1, V[1], c[1])))
    U = r(p[f'{prefix}.fc_offset.weight'])
    o = F.linear(h2, U, p[f'{prefix}.fc_offset.bias'])               # [4P, 2]
    n = o.numel()
    d = o - targets
    loss = torch.where(d.abs() < beta, 0.5 * d * d / beta, d.abs() - 0.5 * beta).sum() * \
        (loss_weight / n)
    # ---------------------------------------------------------------- backward
    g_o = smooth_l1_grad(o, targets, beta, loss_weight / n)           # fp32, not rounded
    grads[f'{prefix}.fc_offset.bias'] = g_o.sum(0)
    grads[f'{prefix}.fc_offset.weight'] = t(g_o).t() @ h2
    g2 = r((h2 > 0) * (t(g_o) @ U))
    grads[f'{prefix}.fcs.1.bias'] = g2.sum(0)
    grads[f'{prefix}.fcs.1.weight'] = g2.t() @ h1
    g1 = r((h1 > 0) * (g2 @ V[1]))
    grads[f'{prefix}.fcs.0.bias'] = g1.sum(0)
    grads[f'{prefix}.fcs.0.weight'] = g1.t() @ h0
    g0 = r((h0 > 0) * (g1 @ V[0]))                                    # [4P, 12544]
    for bi in range(nb):
        seq = acts[bi]
        g = g0[bi * P:(bi + 1) * P].reshape(seq[-1].shape)
        for cidx in range(num_convs - 1, -1, -1):
            a_in = seq[cidx]
            w = r(p[f'{prefix}.expand_convs.{bi}.{cidx}.weight'])
            grads[f'{prefix}.expand_convs.{bi}.{cidx}.bias'] = g.sum((0, 2, 3))
            grads[f'{prefix}.expand_convs.{bi}.{cidx}.weight'] = torch.nn.grad.conv2d_weight(
                a_in, w.shape, g, padding=1)
            if cidx > 0:
                g = r((a_in > 0) * F.conv_transpose2d(g, w, padding=1))
    return loss, grads


def foa_head_backward_teacher_forced(acts, h1, h2, o, targets, p, prefix='roi_head.offset_head',
                                     loss_weight=16.0, beta=1.0):
    """Layer-local check of the FOA head's backward: plain fp32 back-propagation through the head
    using the activations ANOTHER implementation produced (`acts[l]`: [4P,256,7,7] branch-major
    input of conv layer l, acts[-1] the last conv output; h1 / h2: the fc outputs; o: [4P,2]
    predictions) -- identical ReLU masks and identical saved operands on both sides, so what is
    compared is the backward arithmetic itself (weight / bias gradients, mask chain), free of the
    chaotic mask flips that separate two forward passes (see the module docstring).  Weights are
    the TF32-rounded copies the product's tensor cores read.  Runs on the tensors' device."""
    dev = o.device
    nb = 4
    P = o.shape[0] // nb
    W = lambda k: rna(p[k].float().cpu()).to(dev)
    grads = {}
    n = o.numel()
    g_o = smooth_l1_grad(o.float(), targets.to(dev).float(), beta, loss_weight / n)
    U = W(f'{prefix}.fc_offset.weight')
    V = [W(f'{prefix}.fcs.{i}.weight') for i in range(2)]
    h0 = acts[-1].reshape(nb * P, -1)
    grads[f'{prefix}.fc_offset.bias'] = g_o.sum(0)
    grads[f'{prefix}.fc_offset.weight'] = g_o.t() @ h2
    g2 = (h2 > 0) * (g_o @ U)
    grads[f'{prefix}.fcs.1.bias'] = g2.sum(0)
    grads[f'{prefix}.fcs.1.weight'] = g2.t() @ h1
    g1 = (h1 > 0) * (g2 @ V[1])
    grads[f'{prefix}.fcs.0.bias'] = g1.sum(0)
    grads[f'{prefix}.fcs.0.weight'] = g1.t() @ h0
    g0 = ((h0 > 0) * (g1 @ V[0])).reshape(acts[-1].shape)
    num_convs = len(acts) - 1
    for bi in range(nb):
        g = g0[bi * P:(bi + 1) * P]
        for c in range(num_convs - 1, -1, -1):
            a_in = acts[c][bi * P:(bi + 1) * P]
            w = W(f'{prefix}.expand_convs.{bi}.{c}.weight')
            grads[f'{prefix}.expand_convs.{bi}.{c}.bias'] = g.sum((0, 2, 3))
            grads[f'{prefix}.expand_convs.{bi}.{c}.weight'] = torch.nn.grad.conv2d_weight(
                a_in, w.shape, g, padding=1)
            if c > 0:
                g = (a_in > 0) * F.conv_transpose2d(g, w, padding=1)
    return grads
