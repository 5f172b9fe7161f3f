"""Developer tool: time nt_knn on REAL layer-2 features (conv0 output of the bench model) for the NT_KNN_VARIANT set by the caller."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import garment_pattern_estimation_b200 as g
from garment_pattern_estimation_b200 import ops
from garment_pattern_estimation_b200 import configs as om
dev = torch.device('cuda:0')
B, N = 32, 2048
lc = {'loss_components': ['shape', 'loop', 'rotation', 'translation'], 'quality_components': [], 'panel_origin_invariant_loss': False, 'panel_order_inariant_loss': False}
torch.manual_seed(916143406)
model = g.GarmentSegmentPattern3D(dict(om.ATT_DATA_CONFIG), dict(om.ATT_NN_CONFIG), lc).to(dev).train()
pos = torch.randn(B, N, 3, generator=torch.Generator().manual_seed(1234)).to(dev)
with torch.no_grad():
    feats = model.feature_extractor.conv_layers[0](pos.reshape(-1, 3), cloud_shape=(B, N)).contiguous()
def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
ref = None
ms = t(lambda: ops.knn_graph(feats, B, N, 5))
idx = ops.knn_graph(feats, B, N, 5)
print('variant', os.environ.get('NT_KNN_VARIANT', '0'), 'real-features knn150 ms %.3f' % ms, 'checksum', int(idx.long().sum()), ' iid-gaussian ms %.3f' % t(lambda: ops.knn_graph(torch.randn_like(feats), B, N, 5)))
