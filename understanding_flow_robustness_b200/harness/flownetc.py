"""FlowNetC-shaped network used to measure the correlation operator in context (BASELINE config 2/4).

Layer table follows the published FlowNetC architecture as instantiated by the reference
(models/FlowNetC.py:20-49: conv1 7x7/2 64, conv2 5x5/2 128, conv3 5x5/2 256 on both frames, the
21x21 dilation-2 correlation of the two conv3 maps divided by the channel count
(models/submodules.py:124-138), LeakyReLU(0.1), a 1x1 conv_redir to 32 channels, conv3_1..conv6_1, and
the four-level refinement decoder; eval mode returns flow2 * div_flow upsampled x4,
models/FlowNetC.py:193-197).  Random initialisation (Xavier-uniform weights as at :53-62); this is a
throughput / gradient harness, not a port of trained weights.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..spatial_correlation_sampler import spatial_correlation_sample


def correlate(input1, input2):
    """models/submodules.py:124-138: patch 21, dilation_patch 2, collapsed to (B, 441, H, W), / C."""
    # the operator wants dense NCHW (CHECK_CONTIGUOUS in the reference); a channels-last conv stack
    # pays one layout copy here
    out = spatial_correlation_sample(input1.contiguous(), input2.contiguous(), kernel_size=1, patch_size=21,
                                     stride=1, padding=0, dilation_patch=2)
    b, ph, pw, h, w = out.size()
    return out.view(b, ph * pw, h, w) / input1.size(1)


def _conv(cin, cout, k=3, s=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, (k - 1) // 2, bias=True), nn.LeakyReLU(0.1, inplace=True))


def _deconv(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=False), nn.LeakyReLU(0.1, inplace=True))


def _flow(cin):
    return nn.Conv2d(cin, 2, 3, 1, 1, bias=False)


class FlowNetCHarness(nn.Module):
    ENCODER = [("conv1", 3, 64, 7, 2), ("conv2", 64, 128, 5, 2), ("conv3", 128, 256, 5, 2)]
    TRUNK = [("conv3_1", 473, 256, 3, 1), ("conv4", 256, 512, 3, 2), ("conv4_1", 512, 512, 3, 1),
             ("conv5", 512, 512, 3, 2), ("conv5_1", 512, 512, 3, 1), ("conv6", 512, 1024, 3, 2),
             ("conv6_1", 1024, 1024, 3, 1)]

    def __init__(self, div_flow=20.0, corr_fn=correlate, fused_merge=False):
        super().__init__()
        self.div_flow = div_flow
        self.corr_fn = corr_fn
        self.fused_merge = fused_merge   # correlate -> LeakyReLU -> cat as one kernel (merge_block.correlate_merge)
        for name, cin, cout, k, s in self.ENCODER + self.TRUNK:
            setattr(self, name, _conv(cin, cout, k, s))
        self.conv_redir = _conv(256, 32, 1, 1)
        self.deconv5, self.deconv4, self.deconv3, self.deconv2 = (_deconv(1024, 512), _deconv(1026, 256),
                                                                 _deconv(770, 128), _deconv(386, 64))
        self.predict6, self.predict5, self.predict4, self.predict3, self.predict2 = (
            _flow(1024), _flow(1026), _flow(770), _flow(386), _flow(194))
        self.up6, self.up5, self.up4, self.up3 = (nn.ConvTranspose2d(2, 2, 4, 2, 1) for _ in range(4))
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.uniform_(m.bias)
        self.register_buffer("mean", torch.tensor([0.45, 0.432, 0.411]).view(1, 3, 1, 1))

    def features(self, x):
        a1 = self.conv1(x - self.mean)
        a2 = self.conv2(a1)
        return a2, self.conv3(a2)

    def forward(self, img1, img2):
        c2a, c3a = self.features(img1)
        _, c3b = self.features(img2)
        if self.fused_merge:
            from ..merge_block import correlate_merge

            x3 = self.conv3_1(correlate_merge(c3a, c3b, self.conv_redir(c3a), 21, 2, 0.1))
        else:
            corr = F.leaky_relu(self.corr_fn(c3a, c3b), 0.1)
            x3 = self.conv3_1(torch.cat((self.conv_redir(c3a), corr), 1))
        x4 = self.conv4_1(self.conv4(x3))
        x5 = self.conv5_1(self.conv5(x4))
        x6 = self.conv6_1(self.conv6(x5))
        f6 = self.predict6(x6)
        cat5 = torch.cat((x5, self.deconv5(x6), self.up6(f6)), 1)
        f5 = self.predict5(cat5)
        cat4 = torch.cat((x4, self.deconv4(cat5), self.up5(f5)), 1)
        f4 = self.predict4(cat4)
        cat3 = torch.cat((x3, self.deconv3(cat4), self.up4(f4)), 1)
        f3 = self.predict3(cat3)
        cat2 = torch.cat((c2a, self.deconv2(cat3), self.up3(f3)), 1)
        f2 = self.predict2(cat2)
        return F.interpolate(f2 * self.div_flow, scale_factor=4, mode="bilinear", align_corners=False)
