"""BASELINE config 4 (reference-sharded alignment) against the ORACLE, not against a single-GPU run of the same code:
the per-rank work of a 2-rank run is executed rank after rank on one GPU (local buffers stand in for the peers'
symmetric-memory buffers, which is all the DCN epilogue sees of a peer), for both exchange layouts -- every rank holds
every reference's full planes (all-gather layout) and every rank holds every reference's rows of ITS pixel slab -- and
the fused result is compared with matcher oracle -> pre-offset oracle -> DynAgg glue oracle -> plain-C DCN oracle ->
fusion oracle (ref_map_util.py:26-86, corres_generation_arch.py:70-105, ref_mrapa_restoration_arch.py:55-76, :321-335)."""
import pytest
import torch
import torch.nn.functional as F

import mrefsr_b200 as M
import oracle
from mrefsr_b200 import parallel as P
from mrefsr_b200.dcn import dynagg_dcn_forward_into
from oracle.dcn import modulated_deform_conv_c
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('layout', ['full_planes', 'pixel_slabs'])
def test_reference_sharded_alignment_matches_the_oracle(layout):
    n, R, world, h, dg = 1, 4, 2, 16, 8
    scales = ((256, 1), (128, 2), (64, 4))
    g = torch.Generator().manual_seed(5)
    feat_in = torch.randn(n, 256, h, h, generator=g)
    feat_ref = torch.randn(R, n, 256, h, h, generator=g)
    x = {c: torch.randn(R, n, c, h * s, h * s, generator=g) for c, s in scales}
    conv = {c: torch.randn(R, n, 216, h * s, h * s, generator=g) * 0.5 for c, s in scales}
    wgt = {c: torch.randn(c, c, 3, 3, generator=g) * (c * 9) ** -0.5 for c, s in scales}
    bias = {c: torch.randn(c, generator=g) * 0.1 for c, s in scales}
    emb_t = {c: torch.randn(n, c, h * s, h * s, generator=g) * 0.2 for c, s in scales}

    # ---- the sharded computation, rank after rank
    bufs = {}
    for c, s in scales:
        hs = h * s // world
        shape = (n, R, c, hs, h * s) if layout == 'pixel_slabs' else (n, R, c, h * s, h * s)
        bufs[c] = [torch.zeros(shape, device=DEV) for _ in range(world)]
    idx_all = {}
    for rank in range(world):
        lo, hi = P.shard_range(R, rank, world)
        mine = list(range(lo, hi))
        fr = torch.stack([feat_ref[r] for r in mine], 0).flatten(0, 1).to(DEV)
        idx, _ = M.feature_match_index_batched(feat_in.to(DEV), fr, is_norm=True, norm_input=True, normalize_pixels=True,
                                               in_div=1)
        for j, r in enumerate(mine):
            idx_all[r] = idx[j].cpu()
        for c, s in scales:
            xs = torch.stack([x[c][r] for r in mine], 0).flatten(0, 1).to(DEV)
            cs = torch.stack([conv[c][r] for r in mine], 0).flatten(0, 1).to(DEV)
            ptrs = [t.data_ptr() for t in bufs[c]]
            dynagg_dcn_forward_into(xs, cs, idx, s, wgt[c].to(DEV), bias[c].to(DEV), dg, ptrs, len(mine), R, lo,
                                    slab_rows=(h * s // world) if layout == 'pixel_slabs' else 0)
    torch.cuda.synchronize()
    outs = {}
    for c, s in scales:
        if layout == 'pixel_slabs':      # every rank fuses its rows; the rows are then put back together
            parts = []
            for rank in range(world):
                emb = bufs[c][rank].flatten(0, 1)
                parts.append(M.mrapa_attention(P.slab_of(emb_t[c].to(DEV), rank, world), emb, emb.repeat(1, 2, 1, 1), R))
            outs[c] = torch.cat(parts, dim=2)
        else:                            # every rank holds the full planes: any rank's copy gives the result
            assert torch.equal(bufs[c][0], bufs[c][1])
            emb = bufs[c][0].flatten(0, 1)
            outs[c] = M.mrapa_attention(emb_t[c].to(DEV), emb, emb.repeat(1, 2, 1, 1), R)

    # ---- the oracle chain
    a = F.normalize(feat_in[0].reshape(256, -1), dim=0).view(256, h, h)
    for r in range(R):
        q = F.normalize(feat_ref[r, 0].reshape(256, -1), dim=0).view(256, h, h)
        o_idx, _, gap = oracle.feature_match_index_oracle(a, q, is_norm=True, norm_input=True, return_gap=True,
                                                          dtype=torch.float64)
        assert not ((idx_all[r] != o_idx) & (gap >= 1e-5)).any(), 'arg-max mismatch on reference %d' % r
    key = {1: 'relu3_1', 2: 'relu2_1', 4: 'relu1_1'}
    for c, s in scales:
        aligned = []
        for r in range(R):
            pre = oracle.pre_offsets_oracle(idx_all[r])[key[s]][None]
            off, mask = oracle.dynagg_offsets_oracle(conv[c][r], pre, dg)
            aligned.append(modulated_deform_conv_c(x[c][r], off, mask, wgt[c], bias[c], 1, 1, 1, 1, dg))
        emb = torch.stack(aligned, 1).flatten(0, 1)                       # [n*R, C, H, W], references of an image adjacent
        ref = oracle.mrapa_attention_oracle(emb_t[c], emb, emb.repeat(1, 2, 1, 1), R, dtype=torch.float64)
        assert rel_err(outs[c], ref) <= 1e-3, c
