"""FlowNetC body with the merge block on this package's fused kernel (BASELINE config 2 / 4).

Same layers, parameter names and constants as the reference's `models/FlowNetC.py:11-197` (conv1 7x7/2
64, conv2 5x5/2 128, conv3 5x5/2 256 on both frames; `correlate` = the 21x21 dilation-2 correlation of
the conv3 maps / C, `models/submodules.py:124-138`; LeakyReLU(0.1); 1x1 `conv_redir`; conv3_1..conv6_1;
the four-level refinement decoder with biased `predict_flow*` / `deconv*` / `upsampled_flow*`; eval mode
returns `upsample1(flow2 * div_flow)`, :193-197), so a `state_dict` of the reference network loads
unchanged and both produce the same flow (`tests/test_reference_models_gpu.py`).  What differs is only
what SURVEY 8(f) row 2 asks for: with `fused_merge=True` the chain correlate -> /C -> LeakyReLU -> cat
is ONE kernel writing into the concat tensor (`merge_block.correlate_merge`).  The attack benches run
the reference's own file (`harness.reference_models.reference_flownetc`); this class is the same network
with the fused merge block.
"""
import torch
import torch.nn as nn

from ..spatial_correlation_sampler import spatial_correlation_sample


def correlate(input1, input2):
    """models/submodules.py:124-138: patch 21, dilation_patch 2, collapsed to (B, 441, H, W), / C."""
    out = spatial_correlation_sample(input1.contiguous(), input2.contiguous(), kernel_size=1, patch_size=21,
                                     stride=1, padding=0, dilation_patch=2)
    b, ph, pw, h, w = out.size()
    return out.view(b, ph * pw, h, w) / input1.size(1)


def _conv(cin, cout, k=3, s=1):
    # submodules.py:46-58 (batchNorm=False branch)
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, (k - 1) // 2, bias=True), nn.LeakyReLU(0.1, inplace=True))


def _deconv(cin, cout):
    # submodules.py:88-94
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=True), nn.LeakyReLU(0.1, inplace=True))


def _predict_flow(cin):
    # submodules.py:84-85
    return nn.Conv2d(cin, 2, 3, 1, 1, bias=True)


class FlowNetCHarness(nn.Module):
    ENCODER = [("conv1", 3, 64, 7, 2), ("conv2", 64, 128, 5, 2), ("conv3", 128, 256, 5, 2)]
    TRUNK = [("conv3_1", 473, 256, 3, 1), ("conv4", 256, 512, 3, 2), ("conv4_1", 512, 512, 3, 1),
             ("conv5", 512, 512, 3, 2), ("conv5_1", 512, 512, 3, 1), ("conv6", 512, 1024, 3, 2),
             ("conv6_1", 1024, 1024, 3, 1)]
    MEAN = (0.40066648, 0.39482617, 0.3784785)      # FlowNetC.py:73-74 (RGB)

    def __init__(self, div_flow=20.0, corr_fn=correlate, fused_merge=False):
        super().__init__()
        self.div_flow = div_flow
        self.corr_fn = corr_fn
        self.fused_merge = fused_merge   # correlate -> LeakyReLU -> cat as one kernel (merge_block.correlate_merge)
        for name, cin, cout, k, s in self.ENCODER + self.TRUNK:
            setattr(self, name, _conv(cin, cout, k, s))
        self.conv_redir = _conv(256, 32, 1, 1)
        self.corr_activation = nn.LeakyReLU(0.1, inplace=True)
        self.deconv5, self.deconv4, self.deconv3, self.deconv2 = (_deconv(1024, 512), _deconv(1026, 256),
                                                                 _deconv(770, 128), _deconv(386, 64))
        for n, cin in ((6, 1024), (5, 1026), (4, 770), (3, 386), (2, 194)):
            setattr(self, f"predict_flow{n}", _predict_flow(cin))
        for n in (6, 5, 4, 3):
            setattr(self, f"upsampled_flow{n}_to_{n - 1}", nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=True))
        for m in self.modules():                    # FlowNetC.py:53-62
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                if m.bias is not None:
                    nn.init.uniform_(m.bias)
                nn.init.xavier_uniform_(m.weight)
        self.upsample1 = nn.Upsample(scale_factor=4, mode="bilinear")
        self.register_buffer("mean", torch.tensor(self.MEAN, dtype=torch.float64).view(1, 3, 1, 1), persistent=False)

    def features(self, x):
        # normalize_correctly: double mean, cast back (:72-79,92-93; `.float()` there -- the weights' dtype here)
        a1 = self.conv1((x - self.mean).to(self.conv1[0].weight.dtype))
        a2 = self.conv2(a1)
        return a2, self.conv3(a2)

    def forward(self, img1, img2):
        c2a, c3a = self.features(img1)
        _, c3b = self.features(img2)
        if self.fused_merge:
            from ..merge_block import correlate_merge

            x3 = self.conv3_1(correlate_merge(c3a, c3b, self.conv_redir(c3a), 21, 2, 0.1))
        else:
            corr = self.corr_activation(self.corr_fn(c3a, c3b))
            x3 = self.conv3_1(torch.cat((self.conv_redir(c3a), corr), 1))
        x4 = self.conv4_1(self.conv4(x3))
        x5 = self.conv5_1(self.conv5(x4))
        x6 = self.conv6_1(self.conv6(x5))
        f6 = self.predict_flow6(x6)
        cat5 = torch.cat((x5, self.deconv5(x6), self.upsampled_flow6_to_5(f6)), 1)
        f5 = self.predict_flow5(cat5)
        cat4 = torch.cat((x4, self.deconv4(cat5), self.upsampled_flow5_to_4(f5)), 1)
        f4 = self.predict_flow4(cat4)
        cat3 = torch.cat((x3, self.deconv3(cat4), self.upsampled_flow4_to_3(f4)), 1)
        f3 = self.predict_flow3(cat3)
        cat2 = torch.cat((c2a, self.deconv2(cat3), self.upsampled_flow3_to_2(f3)), 1)
        f2 = self.predict_flow2(cat2)
        return self.upsample1(f2 * self.div_flow)
