"""CPU check of the mbarrier protocol of the tcgen05-tail variant of the one-kernel DeepFM forward
(tools/sim/tc_tail_protocol.py models producer / issuer / tensor pipe / gather warps / epilogue warps of
rec_pangu_b200/csrc/deepfm_fused.cu and asserts liveness and operand/accumulator hazards under random interleavings)."""
import importlib.util
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('tc_tail_protocol', os.path.join(ROOT, 'tools', 'sim', 'tc_tail_protocol.py'))
model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(model)


@pytest.mark.parametrize('tiles', [1, 2, 3, 4])
@pytest.mark.parametrize('n_tail', [1, 2, 3])
def test_tc_tail_protocol_is_live_and_hazard_free(tiles, n_tail):
    for nkb in (1, 14):
        for seed in range(8):
            model.Sim(tiles, nkb, n_tail, random.Random(seed + 100 * tiles + 10 * n_tail + nkb)).run()


def test_model_detects_a_missing_drain():
    """The model is only evidence if it can fail: without draining tile t-2's tail before its accumulator buffer is reused
    the issuer and the epilogue wait for each other."""
    class NoDrain(model.Sim):
        def issuer(self):
            g = 0
            for t in range(self.tiles):
                yield lambda t=t: self.tmem_empty[t & 1].done(((t >> 1) & 1) ^ 1)
                for kb in range(self.nkb):
                    s, o = g % model.LB, g % model.OPN
                    yield lambda s=s, g=g: self.full_b[s].done((g // model.LB) & 1)
                    yield lambda o=o, g=g: self.ready_op[o].done((g // model.OPN) & 1)
                    self.pipe += [('main', t, kb, g, s, o), ('commit', self.empty_op[o]), ('commit', self.empty_b[s])]
                    g += 1
                self.pipe.append(('commit', self.tmem_full[t & 1]))
    with pytest.raises(RuntimeError, match='DEADLOCK'):
        NoDrain(3, 2, 2, random.Random(1)).run()


def test_colsum16_butterfly_mapping():
    """numpy restatement of colsum16 (rec_pangu_b200/csrc/tower_tc.cu): after the 4 exchange stages + one xor-16 add, lane l
    holds the sum over the warp's 32 rows of column (l & 15) — the mapping the bias-gradient reductions rely on."""
    import numpy as np
    rng = np.random.default_rng(0)
    V = rng.standard_normal((32, 16))
    v, lanes = V.copy(), np.arange(32)
    off = 8
    while off > 0:
        up = (lanes & off) != 0
        new = v.copy()
        for i in range(off):
            send = np.where(up, v[:, i], v[:, i + off])
            new[:, i] = np.where(up, v[:, i + off], v[:, i]) + send[lanes ^ off]
        v, off = new, off // 2
    res = v[:, 0] + v[lanes ^ 16, 0]
    assert np.allclose(res, V.sum(axis=0)[lanes & 15])


# ---------------------------------------------------------------- cin_wgrad_tc_kernel (tools/sim/cin_wgrad_protocol.py)
spec2 = importlib.util.spec_from_file_location('cin_wgrad_protocol', os.path.join(ROOT, 'tools', 'sim', 'cin_wgrad_protocol.py'))
wg = importlib.util.module_from_spec(spec2)
spec2.loader.exec_module(wg)


@pytest.mark.parametrize('nkb', [0, 1, 2, 3, 4, 5, 6, 7, 13, 40])
def test_cin_wgrad_protocol_is_live_and_hazard_free(nkb):
    """Two cp.async loader warps (4 groups in flight over 6 stages), 8 software-pipelined operand warps, MMA issuer, asynchronous
    tensor pipe and copy completion of cin_wgrad_tc_kernel: no deadlock, every read sees completely landed data of the right
    k-block, for any number of k-blocks relative to the ring depths."""
    for seed in range(12):
        wg.Sim(nkb, random.Random(1000 * nkb + seed)).run()


def _fails(make, seeds=40):
    for seed in range(seeds):
        try:
            make(random.Random(seed)).run()
        except (AssertionError, RuntimeError):
            return True
    return False


def test_cin_wgrad_model_detects_protocol_mutations():
    """The model is only evidence if it can fail: (a) waiting for one cp.async group too few lets the lo pass / the operand
    warps read a stage that has not landed; (b) releasing the G / X0 stage before the last read lets the loader refill it under
    a reader; (c) more groups in flight than stages: the prologue waits for a stage nobody can have consumed yet."""
    def slack(rng):
        s = wg.Sim(13, rng)
        s.wait_slack = 1
        return s

    def early(rng):
        s = wg.Sim(13, rng)
        s.early_release = True
        return s
    assert _fails(slack)
    assert _fails(early)
    assert _fails(lambda rng: wg.Sim(13, rng, bst=4, ahead=5))
    assert not _fails(lambda rng: wg.Sim(13, rng, bst=4, ahead=4), seeds=10)      # as many as stages is legal (no slack, no hazard)
    assert not _fails(lambda rng: wg.Sim(13, rng), seeds=10)


# ---------------------------------------------------------------- cin_fwd2_tc_kernel / cin_bwd_tc_kernel (tools/sim/cin_tile_protocol.py)
spec3 = importlib.util.spec_from_file_location('cin_tile_protocol', os.path.join(ROOT, 'tools', 'sim', 'cin_tile_protocol.py'))
tp = importlib.util.module_from_spec(spec3)
spec3.loader.exec_module(tp)


@pytest.mark.parametrize('ntiles', [2, 3, 4])
def test_cin_tile_protocol_is_live_and_hazard_free(ntiles):
    """Operand double buffer + accumulator ping-pong with a counter that runs across tiles (NTILES = 3: cin_fwd2 at M = 26;
    2 / 4: the other instantiations), four operand and four epilogue warps, asynchronous tensor pipe."""
    for tiles in (0, 1, 2, 3, 5, 8):
        for seed in range(10):
            tp.Sim(tiles, ntiles, random.Random(97 * tiles + 7 * ntiles + seed)).run()


def test_cin_tile_model_detects_an_early_operand_release():
    def early(rng):
        s = tp.Sim(6, 3, rng)
        s.early_a_release = True
        return s
    assert _fails(early, seeds=60)
    assert not _fails(lambda rng: tp.Sim(6, 3, rng), seeds=10)
