"""Time the torch/cuDNN template encoder (2 images, fp32, TF32 off) in a few configurations
that keep fp32 arithmetic: plain, channels_last, cudnn.benchmark, BN folded into the convs,
CUDA graph.  Prints ms per call and the max deviation of each variant from the plain output."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from nope_b200 import synth_weights as weights
from nope_b200.encoder import FeatureExtractor

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
enc_sd = {k: v for k, v in weights.make_encoder_state_dict(0).items()}
fe = FeatureExtractor(descriptor_size=8).to(dev)
fe.load_state_dict(enc_sd)
x = (torch.rand(2, 3, 256, 256, device=dev) * 2 - 1)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ref = fe.encode_image(x)
    print("plain            %.3f ms" % timeit(lambda: fe.encode_image(x)))
    torch.backends.cudnn.benchmark = True
    print("cudnn.benchmark  %.3f ms" % timeit(lambda: fe.encode_image(x)),
          "dev", float((fe.encode_image(x) - ref).abs().max()))
    fe_cl = FeatureExtractor(descriptor_size=8).to(dev)
    fe_cl.load_state_dict(enc_sd)
    fe_cl = fe_cl.to(memory_format=torch.channels_last)
    xc = x.contiguous(memory_format=torch.channels_last)
    print("channels_last    %.3f ms" % timeit(lambda: fe_cl.encode_image(xc)),
          "dev", float((fe_cl.encode_image(xc) - ref).abs().max()))

    # fold eval-mode BatchNorm into the preceding conv (exact up to fp32 rounding)
    def fold(conv, bn):
        w = conv.weight
        s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        new = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride,
                        conv.padding, bias=True).to(dev)
        new.weight.copy_(w * s[:, None, None, None])
        new.bias.copy_(bn.bias - bn.running_mean * s)
        return new

    fe_f = FeatureExtractor(descriptor_size=8).to(dev)
    fe_f.load_state_dict(enc_sd)
    bb = fe_f.backbone
    bb.conv1, bb.bn1 = fold(bb.conv1, bb.bn1), nn.Identity()
    for layer in (bb.layer1, bb.layer2, bb.layer3, bb.layer4):
        for blk in layer:
            blk.conv1, blk.bn1 = fold(blk.conv1, blk.bn1), nn.Identity()
            blk.conv2, blk.bn2 = fold(blk.conv2, blk.bn2), nn.Identity()
            blk.conv3, blk.bn3 = fold(blk.conv3, blk.bn3), nn.Identity()
            if blk.downsample is not None:
                blk.downsample = nn.Sequential(fold(blk.downsample[0], blk.downsample[1]))
    print("bn folded        %.3f ms" % timeit(lambda: fe_f.encode_image(x)),
          "dev", float((fe_f.encode_image(x) - ref).abs().max()), "ref scale", float(ref.abs().max()))
    fe_fc = fe_f.to(memory_format=torch.channels_last)
    print("folded + NHWC    %.3f ms" % timeit(lambda: fe_fc.encode_image(xc)),
          "dev", float((fe_fc.encode_image(xc) - ref).abs().max()))

    # CUDA graph of the folded encoder
    static_x = xc.clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fe_fc.encode_image(static_x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_y = fe_fc.encode_image(static_x)
    print("folded+NHWC+graph %.3f ms" % timeit(lambda: g.replay()),
          "dev", float((static_y - ref).abs().max()))
    # TF32 for reference (changes numerics)
    torch.backends.cudnn.allow_tf32 = True
    fe_t = FeatureExtractor(descriptor_size=8).to(dev)
    fe_t.load_state_dict(enc_sd)

    def tf32():
        return fe_t.projector(fe_t.backbone(x))
    print("tf32 (numerics!) %.3f ms" % timeit(tf32), "dev", float((tf32() - ref).abs().max()))
    with torch.autocast("cuda", dtype=torch.float16):
        print("fp16 autocast    %.3f ms" % timeit(tf32), "dev", float((tf32().float() - ref).abs().max()))
