"""
A random-init torch STAND-IN for the RetinaNet-3D detector.  The CNN is out of scope
(/root/reference/keras_retinanet_3D/models/retinanet.py:22-340 needs Keras/TF weights that cannot exist here); this
module only produces head tensors of the right shape, layout and rough statistics -- per pyramid level P3..P7, per
cell, 12 anchors x (12 box/key-point offsets, 3 dimension offsets, 8 orientation-class scores) in the reference's
anchor order (utils/anchors.py) -- so that the driver and the device pipeline after the CNN can run end to end.
Its outputs mean nothing; parity is defined on whatever heads it (or any real detector) hands over.
"""
import torch

__all__ = ['StandInDetector']

ANCHORS_PER_CELL = 12


class StandInDetector(torch.nn.Module):
    def __init__(self, seed=0, score_bias=-7.0, levels=(3, 4, 5, 6, 7)):
        super().__init__()
        self.levels = tuple(levels)
        g = torch.Generator().manual_seed(int(seed))
        k = ANCHORS_PER_CELL
        self.stem = torch.nn.Conv2d(3, 16, 3, padding=1)
        self.reg = torch.nn.Conv2d(16, k * 12, 3, padding=1)
        self.dim = torch.nn.Conv2d(16, k * 3, 3, padding=1)
        self.cls = torch.nn.Conv2d(16, k * 8, 3, padding=1)
        with torch.no_grad():
            for m, gain in ((self.stem, 0.03), (self.reg, 1.0 / 12), (self.dim, 1.0 / 12), (self.cls, 1.5 / 12)):
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * gain)
                m.bias.zero_()
            self.cls.bias.fill_(float(score_bias))       # few anchors above the 0.05 score threshold

    @torch.no_grad()
    def forward(self, image):
        """image (B, rows, cols, 3) float32, preprocessed BGR.  Returns regression (B, A, 12),
        regression_dim (B, A, 3), classification (B, A, 8) with A = sum over levels of cells * 12."""
        x = image.permute(0, 3, 1, 2).contiguous()
        B, _, rows, cols = x.shape
        regs, dims, clss = [], [], []
        for p in self.levels:
            h, w = (rows + 2 ** p - 1) // 2 ** p, (cols + 2 ** p - 1) // 2 ** p        # utils/anchors.py:140-152
            f = torch.tanh(self.stem(torch.nn.functional.adaptive_avg_pool2d(x, (h, w))))
            # (B, k * c, h, w) -> (B, h * w * k, c): cells row-major, the 12 anchors of a cell contiguous
            to_rows = lambda t, c: t.permute(0, 2, 3, 1).reshape(B, h * w * ANCHORS_PER_CELL, c)  # noqa: E731
            regs.append(to_rows(self.reg(f), 12))
            dims.append(to_rows(self.dim(f), 3))
            clss.append(torch.sigmoid(to_rows(self.cls(f), 8)))
        return torch.cat(regs, 1).contiguous(), torch.cat(dims, 1).contiguous(), torch.cat(clss, 1).contiguous()
