"""Drop-in replacement for the reference's `model.py` (module names, constructor
signatures, forward signatures and state_dict keys are the reference's:
model.py:6-210), with the decoder tail of every stage running in the fused
sm_100a kernels of libpwr_b200.so.

    from pixelwiseregression_b200.model import PixelwiseRegression   # instead of `from model import ...`

What is unchanged, on purpose: the hourglass / conv backbone stays on cuDNN
("the unchanged reference", north_star); its layer layout inside every
`torch.nn.Sequential` is kept index-for-index so released checkpoints load
(utils.py:309-314) — 344 state_dict entries for the NYU configuration.

What changed: `PredictionBlock.forward` no longer runs softmax / mul / sum /
cat as ~16 ATen kernels (model.py:83-95, 123-130, 151); it hands the two conv
outputs to `ops.fused_decoder` (one forward kernel, one backward kernel).
`PlaneRegression.forward` and `DepthRegression.forward` stay individually
callable with their reference signatures and also run on the fused kernels.
CUDA only: there is no CPU fallback.
"""
import torch

from . import ops


def com_filter(size_u, size_v):
    """utils.py:24-35 as a [2, size_v, size_u] float32 tensor (channel 0: U,
    channel 1: V), built in float64 like the NumPy original."""
    u = (torch.arange(size_u, dtype=torch.float64) - size_u // 2) / (size_u - 1)
    v = (torch.arange(size_v, dtype=torch.float64) - size_v // 2) / (size_v - 1)
    return torch.stack([u.view(1, -1).expand(size_v, size_u), v.view(-1, 1).expand(size_v, size_u)]).float().contiguous()


def xavier_weights_init(m):
    """utils.py:339-342."""
    if isinstance(m, torch.nn.Conv2d):
        torch.nn.init.xavier_normal_(m.weight.data)


def _norm_relu_conv(norm, inplace, c_in, c_out, k):
    return [norm(c_in, affine=True), torch.nn.ReLU(inplace), torch.nn.Conv2d(c_in, c_out, k, stride=1, padding=k // 2)]


def _conv_norm_relu(norm, inplace, c_in, c_out, k, stride=1):
    return [torch.nn.Conv2d(c_in, c_out, k, stride=stride, padding=k // 2), norm(c_out, affine=True),
            torch.nn.ReLU(inplace)]


def _regression_head(features, joints, kernel_size, norm, inplace):
    """Three conv-norm-relu blocks and a projection to `joints` maps
    (model.py:54-65 and :103-114 share this layout)."""
    layers = []
    for _ in range(3):
        layers += _conv_norm_relu(norm, inplace, features, features, kernel_size)
    layers.append(torch.nn.Conv2d(features, joints, kernel_size, stride=1, padding=kernel_size // 2))
    return torch.nn.Sequential(*layers)


_NORMS = {'batch': torch.nn.BatchNorm2d, 'instance': torch.nn.InstanceNorm2d}


def _stem(features, kernel_size, norm):
    """1 -> 32 -> 64 -> ... -> features channels at 128x128, then a stride-2 conv to 64x64
    (model.py:164-186 and :268-288 share this layout)."""
    layers = _conv_norm_relu(norm, True, 1, 32, kernel_size)
    width = 32
    while width < features:
        nxt = min(2 * width, features)
        layers += _conv_norm_relu(norm, True, width, nxt, kernel_size)
        width = nxt
    layers += _conv_norm_relu(norm, True, features, features, kernel_size, stride=2)
    return torch.nn.Sequential(*layers)


class ResBlock(torch.nn.Module):
    """Pre-activation bottleneck, model.py:6-23."""

    def __init__(self, features, kernel_size=3, norm=torch.nn.BatchNorm2d, inplace=True):
        super().__init__()
        half = features // 2
        self.conv = torch.nn.Sequential(*(_norm_relu_conv(norm, inplace, features, half, 1)
                                          + _norm_relu_conv(norm, inplace, half, half, kernel_size)
                                          + _norm_relu_conv(norm, inplace, half, features, 1)))

    def forward(self, x):
        return x + self.conv(x)


class Hourglass(torch.nn.Module):
    """Recursive hourglass, model.py:25-47 (max-pool down, nearest up)."""

    def __init__(self, features, level=4, kernel_size=3, norm=torch.nn.BatchNorm2d):
        super().__init__()
        self.input_conv = ResBlock(features, kernel_size=kernel_size, norm=norm)
        self.down_sample = torch.nn.MaxPool2d(2, stride=2)
        if level > 0:
            self.inner = Hourglass(features, level - 1, kernel_size=kernel_size, norm=norm)
        else:
            self.inner = ResBlock(features, kernel_size=kernel_size, norm=norm)
        self.output_conv = ResBlock(features, kernel_size=kernel_size, norm=norm)

    def forward(self, x):
        x = self.input_conv(x)
        h = self.output_conv(self.inner(self.down_sample(x)))
        return torch.nn.functional.interpolate(h, size=x.shape[2:]) + x


class PlaneRegression(torch.nn.Module):
    """model.py:49-97.  forward(f) -> (heatmaps [B,J,H,W], plane_coordinates [B,J,2])."""

    def __init__(self, features, joints, label_size, kernel_size=3, norm=torch.nn.BatchNorm2d, inplace=True,
                 normalization_method='softmax'):
        super().__init__()
        self.normalization_method = normalization_method
        self.conv = _regression_head(features, joints, kernel_size, norm, inplace)
        # kept for state_dict compatibility; the kernels synthesise U, V from pixel indices
        self.register_buffer('filter', com_filter(label_size, label_size))
        if normalization_method == 'softmax':
            self.register_parameter('w', torch.nn.Parameter(torch.ones(joints, 1)))

    @property
    def method(self):
        return 'softmax' if self.normalization_method == 'softmax' else 'sum'

    @property
    def temperature(self):
        return self.w if self.normalization_method == 'softmax' else None

    def forward(self, f):
        z = self.conv(f)
        if torch.is_grad_enabled() and (z.requires_grad or self.temperature is not None):
            return ops.PlaneFunction.apply(z, self.temperature, self.method)
        H, uvd, _, _ = ops.decoder_forward_raw(z, self.temperature, None, None, None, self.method, want_stats=False)
        return H, uvd[:, :, :2].contiguous()


class DepthRegression(torch.nn.Module):
    """model.py:99-132.  forward(f, heatmaps, label_img, mask) -> (depthmaps, depth_coordinates [B,J,1])."""

    def __init__(self, features, joints, kernel_size=3, norm=torch.nn.BatchNorm2d, inplace=True):
        super().__init__()
        self.conv = _regression_head(features, joints, kernel_size, norm, inplace)

    def forward(self, f, heatmaps, label_img, mask):
        depthmaps = self.conv(f)
        return depthmaps, ops.DepthFunction.apply(depthmaps, heatmaps, label_img, mask)


class PredictionBlock(torch.nn.Module):
    """One stage, model.py:134-151.  forward(x, label_img, mask) -> (f, heatmaps, depthmaps, uvd [B,J,3])."""

    def __init__(self, in_dim, joints, label_size=64, features=256, level=4, kernel_size=3,
                 norm=torch.nn.BatchNorm2d, heatmap_method='softmax'):
        super().__init__()
        self.conv = torch.nn.Conv2d(in_dim, features, 1, stride=1, padding=0)
        self.hourglass = Hourglass(features, level, norm=norm)
        self.plane_regression = PlaneRegression(features, joints, label_size, kernel_size=kernel_size, norm=norm,
                                                normalization_method=heatmap_method)
        self.depth_regression = DepthRegression(features, joints, kernel_size=kernel_size, norm=norm)

    def features_and_logits(self, x):
        f = self.hourglass(self.conv(x))
        return f, self.plane_regression.conv(f), self.depth_regression.conv(f)

    def forward(self, x, label_img, mask):
        f, z, d_raw = self.features_and_logits(x)
        plane = self.plane_regression
        heatmaps, depthmaps, uvd = ops.fused_decoder(z, plane.temperature, d_raw, label_img, mask, plane.method)
        return f, heatmaps, depthmaps, uvd


class PixelwiseRegression(torch.nn.Module):
    """model.py:153-210.  forward(img [B,1,128,128], label_img [B,1,64,64], mask [B,1,64,64])
    -> list over stages of (heatmaps, depthmaps, uvd)."""

    def __init__(self, joints, stage=2, label_size=64, features=256, level=4, kernel_size=3, norm_method='batch',
                 heatmap_method='softmax'):
        super().__init__()
        norm = _NORMS[norm_method]
        self.conv = _stem(features, kernel_size, norm)
        concat_dim = 2 * joints + 1
        self.stages = torch.nn.ModuleList([
            PredictionBlock(features if i == 0 else concat_dim, joints, label_size, features, level,
                            kernel_size=kernel_size, heatmap_method=heatmap_method, norm=norm)
            for i in range(stage)])
        self.apply(xavier_weights_init)

    def forward(self, img, label_img, mask):
        f = self.conv(img)
        results = []
        for stage in self.stages:
            f, heatmaps, depthmaps, uvd = stage(f, label_img, mask)
            results.append((heatmaps, depthmaps, uvd))
            f = torch.cat([heatmaps, depthmaps, label_img], dim=1)
        return results

    def forward_loss(self, img, label_img, mask, uvd, heatmaps, depthmaps, alpha=1.0, lambda_h=1.0, lambda_d=0.01,
                     n_mean=0):
        """Fused criterion: the training forward plus the loss of train.py:194-205 with
        the loss arithmetic inside the decoder kernels.  Returns (loss, every_loss, uvds):
        `every_loss[i]` is a [3] tensor (heatmap_loss, depthmap_loss, uvd_loss) of stage i
        (what train.py logs, :296-310) and `uvds[i]` the decoded coordinates.  The last
        stage runs forward and backward+loss back to back (ops.fused_decoder_loss).
        `heatmaps` may be the sparse `taps` tensor of sfr.build_sfr(targets="sparse")
        (`depthmaps` is then ignored): the loss kernels evaluate the targets on the fly.
        `n_mean`: the B*J of the loss means (0 = this batch; see ops.fused_decoder_loss)."""
        f = self.conv(img)
        loss = 0
        every_loss, uvds = [], []
        last = len(self.stages) - 1
        for i, stage in enumerate(self.stages):
            f, z, d_raw = stage.features_and_logits(f)
            plane = stage.plane_regression
            if i < last:
                H, D, uvd_i, stage_loss, terms = ops.fused_decoder_with_loss(
                    z, plane.temperature, d_raw, label_img, mask, heatmaps, depthmaps, uvd, plane.method, alpha,
                    lambda_h, lambda_d, n_mean)
                f = torch.cat([H, D, label_img], dim=1)
            else:
                stage_loss, terms, uvd_i = ops.fused_decoder_loss(
                    z, plane.temperature, d_raw, label_img, mask, heatmaps, depthmaps, uvd, plane.method, alpha,
                    lambda_h, lambda_d, store_heat=False, n_mean=n_mean)[:3]
            loss = loss + stage_loss
            every_loss.append(terms)
            uvds.append(uvd_i.detach())
        return loss, every_loss, uvds


class FullRegressionBlock(torch.nn.Module):
    """Ablation baseline, model.py:215-257: hourglass features -> three stride-2 convs -> MLP -> [B,J,3].
    No decoder (nothing of the hot path runs here); kept so that `train_fullregression.py` /
    `test_fullregression.py` find the names they import and released ablation checkpoints load
    (same Sequential layouts, same state_dict keys)."""

    def __init__(self, in_dim, joints, label_size=64, features=256, level=4, norm=torch.nn.BatchNorm2d):
        super().__init__()
        self.conv = torch.nn.Conv2d(in_dim, features, 1, stride=1, padding=0)
        self.hourglass = Hourglass(features, level, norm=norm)
        self.flatten_dim = label_size ** 2 * features // 64
        self.joints = joints
        down = []
        for _ in range(3):
            down += _conv_norm_relu(norm, True, features, features, 3, stride=2)
        self.downsampling = torch.nn.Sequential(*down)
        self.regression = torch.nn.Sequential(torch.nn.Linear(self.flatten_dim, 1024), torch.nn.ReLU(True),
                                              torch.nn.Linear(1024, 1024), torch.nn.ReLU(True),
                                              torch.nn.Linear(1024, joints * 3))

    def forward(self, x, label_img, mask):
        f = self.hourglass(self.conv(x))
        coordinates = self.regression(self.downsampling(f).view(-1, self.flatten_dim))
        return f, coordinates.view(-1, self.joints, 3)


class FullRegression(torch.nn.Module):
    """Ablation baseline, model.py:259-309.  forward(img, label_img, mask) -> list over stages of uvd [B,J,3]."""

    def __init__(self, joints, stage=2, label_size=64, features=256, level=4, norm_method='batch'):
        super().__init__()
        norm = _NORMS[norm_method]
        self.conv = _stem(features, 3, norm)
        self.stages = torch.nn.ModuleList([
            FullRegressionBlock(features if i == 0 else features + 1, joints, label_size, features, norm=norm)
            for i in range(stage)])
        self.apply(xavier_weights_init)

    def forward(self, img, label_img, mask):
        f = self.conv(img)
        results = []
        for stage in self.stages:
            f, uvd = stage(f, label_img, mask)
            results.append(uvd)
            f = torch.cat([f, label_img], dim=1)
        return results


__all__ = ["ResBlock", "Hourglass", "PlaneRegression", "DepthRegression", "PredictionBlock", "PixelwiseRegression",
           "FullRegressionBlock", "FullRegression", "com_filter", "xavier_weights_init"]
