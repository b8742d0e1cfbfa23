"""SparseConvTensor: the container det3d's backbone/neck exchange (spconv.pytorch.SparseConvTensor,
constructed at det3d/ops/pillar_ops/pillar_modules.py:74; consumers use .features, .indices,
.spatial_shape, .batch_size, .replace_feature() and .dense(): backbones/base.py:9-10,139-140,
PillarResNet.py:139, necks/rpn.py:196-199).

Unlike spconv's, the row count lives on the device: buffers are sized by a host capacity and
`table.num` holds the true count, so a forward pass never has to synchronise.  `.features` /
`.indices` give the exact-length views the reference API promises (one host sync, cached)."""
import torch

from . import ops


class SparseConvTensor:
    def __init__(self, feat, table, spatial_shape=None, batch_size=None):
        self.feat = feat              # (cap, C) channels-last, f32 or bf16
        self.table = table            # ops.RankTable (coords, occupancy, count)
        self.spatial_shape = list(spatial_shape) if spatial_shape is not None else [table.H, table.W]
        self.batch_size = batch_size if batch_size is not None else table.B
        self._count = None

    # --- reference-compatible accessors (host sync) ---
    def _n(self):
        if self._count is None:
            self._count = self.table.count()
        return self._count

    @property
    def features(self):
        return self.feat[: self._n()]

    @property
    def indices(self):
        return self.table.coords[: self._n()]

    def replace_feature(self, feat):
        """shares indices / rulebook cache, like spconv's replace_feature."""
        if feat.shape[0] != self.feat.shape[0]:
            n = self._n()
            buf = torch.zeros(self.table.cap, feat.shape[1], dtype=feat.dtype, device=feat.device)
            buf[:n] = feat[:n]
            feat = buf
        out = SparseConvTensor(feat, self.table, self.spatial_shape, self.batch_size)
        out._count = self._count
        return out

    def dense_nhwc(self, out=None, out_coff=0, padded=False):
        """(B*H*W, C) channels-last densify (zero elsewhere); padded: the zero-bordered (H+2,W+2) map."""
        return ops.sparse_to_dense(self.feat, self.table, self.feat.shape[1], out=out, out_coff=out_coff,
                                   padded=padded)

    def dense(self, channels_first=True):
        """(B,C,H,W) tensor as spconv's .dense(); storage is NHWC (a channels_last NCHW view)."""
        d = self.dense_nhwc().view(self.batch_size, self.table.H, self.table.W, -1)
        return d.permute(0, 3, 1, 2) if channels_first else d
