"""Oracle (test infrastructure): YOLOv5 Detect decode.  Reference: basics/models/model.py:48-70."""
import torch


def detect_decode(raw, anchors_px, stride, dtype=torch.float64):
    """raw: [B, na*no, ny, nx] output of the 1x1 conv (model.py:53).

    Returns (z [B, na*ny*nx, no], x_perm [B, na, ny, nx, no]):
      x_perm = raw viewed [B,na,no,ny,nx] and permuted to channels-last (model.py:55);
      y = sigmoid(x_perm); xy = (2y - 0.5 + grid) * stride; wh = (2y)^2 * anchor (model.py:61-63);
      rows of z ordered (anchor, y, x), x fastest; grid = (x index, y index) (model.py:67-70).
    ``anchors_px`` [na, 2] is Detect.anchor_grid (pixels).
    """
    B, ch, ny, nx = raw.shape
    na = anchors_px.shape[0]
    no = ch // na
    xp = raw.reshape(B, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
    y = torch.sigmoid(xp.to(dtype))
    gx = torch.arange(nx, dtype=dtype).reshape(1, 1, 1, nx)
    gy = torch.arange(ny, dtype=dtype).reshape(1, 1, ny, 1)
    out = y.clone()
    out[..., 0] = (y[..., 0] * 2.0 - 0.5 + gx) * stride
    out[..., 1] = (y[..., 1] * 2.0 - 0.5 + gy) * stride
    a = anchors_px.to(dtype).reshape(1, na, 1, 1, 2)
    out[..., 2:4] = (y[..., 2:4] * 2.0) ** 2 * a
    return out.reshape(B, na * ny * nx, no), xp
