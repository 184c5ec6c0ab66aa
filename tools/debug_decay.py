"""Debug aid: replay test_decay_until_removed and print where GPU and oracle feature weights differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import scenes as S
from tests.parity_utils import Pair, make_params, orbit_frames, gpu_blocks

mp, op = make_params(workspace=S.WS_CUBE_STACKING, decay=0.5)
pair = Pair(0.02, 16, mp, op)
for i, T, K, depth, feat in orbit_frames(1, 64, 64, 16, S.S_TABLE):
    pair.depth(depth, T, K)
    pair.features(feat, T, K)
for _ in range(16):
    pair.decay()
print('blocks after decay', pair.gpu.tsdf_layer_view(0).num_blocks(), pair.gpu.feature_layer_view(0).num_blocks())
for i, T, K, depth, feat in orbit_frames(2, 64, 64, 16, S.S_TABLE, seed0=50):
    pair.depth(depth, T, K)
    g, c = pair.last_block_list(0)
    print('frame', i, 'view lists equal', np.array_equal(g, c), len(g))
    pair.features(feat, T, K)
    g, c = pair.last_block_list(1)
    print('  band lists equal', np.array_equal(g, c), len(g), len(c))
    gs, cs = pair.synthetic_depth()
    print('  synth equal', np.array_equal(gs.view(np.uint32), cs.view(np.uint32)))
    gi, gd = gpu_blocks(pair.gpu.feature_layer_view(0))
    ci, cd = pair.cpu.all_blocks(1)
    print('  feature block sets equal', np.array_equal(gi, ci), len(gi))
    g16, c16 = gd.view(np.uint16), cd.view(np.uint16)
    dw = g16[..., -1] != c16[..., -1]
    print('  weight mismatches', int(dw.sum()), 'gpu zero where oracle nonzero', int(((g16[..., -1] == 0) & dw).sum()),
          'gpu nonzero where oracle zero', int(((c16[..., -1] == 0) & dw).sum()))
    if dw.any():
        blocks = np.unique(np.nonzero(dw)[0])
        print('  blocks with mismatches', len(blocks), 'of', len(gi), blocks[:10])
        b = blocks[0]
        print('   per-block mismatch count', [(int(x), int(dw[x].sum())) for x in blocks[:10]])
        df = (g16[..., :-1] != c16[..., :-1]).any(-1)
        print('  feature mismatching voxels', int(df.sum()))
    print('  counters', {k: (pair.gpu.counters(0)[k], pair.cpu.counters()[k]) for k in ('feature_voxels_updated', 'feature_blocks_allocated', 'feature_band_blocks')})
