"""CPU: host-side logic -- state_dict layout, config mirror, op programs (through the reference interpreter,
with the real memory plan), frame sharding.  No GPU, no compute calls into the shared library."""
import copy
import os
import re

import numpy as np
import pytest
import torch

import capf_oracle
import interp
import protocol
from capf_b200 import lib, program
from capf_b200.mvn.utils import cfg as cfgmod
from conftest import build_case_model, golden_cases, load_golden, rel_l2, ROOT


@pytest.mark.parametrize("backbone", ["hrnet_32", "hrnet_48", "cpn"])
def test_state_dict_layout_matches_reference(backbone, manifest):
    m, _, _ = build_case_model(backbone, 0)
    ours = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert sorted(ours) == sorted(manifest[backbone])          # same keys, same shapes (train.py:307-314 strict=True)
    assert all(not p.requires_grad for p in m.backbone.parameters())      # conpose.py:22-25
    assert all(p.requires_grad for p in m.volume_net.parameters())


def test_default_init_matches_reference_quirks():
    """pose_dformer.py:103-113,184: zero attention/offset weights, directional offset bias, zero pos-embed."""
    import capf_b200
    m = capf_b200.CA_PF(capf_b200.make_config("hrnet_32"))
    cb = m.volume_net.context_blocks[0]
    assert cb.attention_weights.weight.abs().sum() == 0 and cb.sampling_offsets.weight.abs().sum() == 0
    assert m.volume_net.Spatial_pos_embed.abs().sum() == 0
    b = cb.sampling_offsets.bias.view(4, 4, 2)
    assert torch.allclose(b[0, :, 0], torch.tensor([0.01, 0.02, 0.03, 0.04])) and b[0, :, 1].abs().max() < 1e-8
    assert m.volume_net.head[0].eps == 1e-5 and m.volume_net.res_blocks[0].norm1.eps == 1e-6


def test_config_mirror(tmp_path):
    c = copy.deepcopy(cfgmod.config)
    assert c.model.backbone.STAGE3.NUM_CHANNELS == [32, 64, 128] and c["model"]["poseformer"]["levels"] == 4
    y = tmp_path / "ok.yaml"
    y.write_text("model:\n  backbone:\n    type: cpn\n    fix_weights: true\n  poseformer:\n    embed_dim_ratio: 128\n")
    saved = copy.deepcopy(dict(cfgmod.config))
    try:
        cfgmod.update_config(str(y))
        assert cfgmod.config.model.backbone.type == "cpn" and cfgmod.config.model.backbone.fix_weights is True
        bad = tmp_path / "bad.yaml"
        bad.write_text("model:\n  not_a_key: 1\n")
        with pytest.raises(ValueError, match="not exist in cfg.py"):       # cfg.py:173-174
            cfgmod.update_config(str(bad))
    finally:
        cfgmod.config.clear()
        cfgmod.config.update(cfgmod.AttrDict(saved))
    ref_yaml = "/root/reference/ContextPose/experiments/human36m/human36m.yaml"
    if os.path.isfile(ref_yaml):                                           # the reference's own YAML loads unchanged
        import capf_b200
        c = capf_b200.make_config("hrnet_48", ref_yaml)
        assert c.model.backbone.STAGE4.NUM_CHANNELS == [48, 96, 192, 384] and c.model.poseformer.base_dim == 48
        assert c.train.batch_size == 512 and c.model.backbone.fix_weights is True


CASES = [c for c in golden_cases() if c["name"] in ("hrnet32_b2_128x96", "hrnet48_b1_384x288", "cpn_b2_256x256")]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_program_matches_golden(case, precision):
    """The op program (as the GPU will run it, including buffer aliasing) reproduces the reference output."""
    g = load_golden(case["name"])
    m, w, cfg = build_case_model(case["backbone"], case["weight_seed"])
    B, H, W = case["B"], case["H"], case["W"]
    images, kp2d, crop = protocol.make_inputs(B, H, W, case["input_seed"])
    prog = program.build_forward_program(case["backbone"], getattr(m.backbone, "cfg", None), m._pf_cfg,
                                         {k: tuple(v.shape) for k, v in w.items()}, B, H, W, precision, debug_records=True)
    it = interp.Interp(prog, w)
    it.t(prog.inputs["images"]).copy_(images)
    it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
    it.t(prog.inputs["ref"]).copy_(torch.from_numpy(g["crop_after"]).reshape(-1, 2))
    it.run()
    out = it.t(prog.outputs["out"]).view(B, 1, 17, 3)
    tol = 1e-5 if precision == "fp32" else 2.5e-3      # fp16 storage: ~7e-4 HRNet, ~1.3e-3 CPN (documented in DESIGN.md)
    assert rel_l2(out, g["out"]) < tol
    for l, f in enumerate(prog.feature_maps):
        nchw = it.t(f).float().permute(0, 3, 1, 2).reshape(-1)
        assert rel_l2(nchw[torch.from_numpy(g[f"feat{l}_idx"])], g[f"feat{l}_val"]) < (1e-5 if precision == "fp32" else 3e-3)


@pytest.mark.parametrize("name", ["hrnet32_b2_128x96", "cpn_b2_256x256"])
def test_split_operand_program_matches_golden(name):
    """precision="bf16x3": fp32 storage, every conv / Linear as hi*Wh + lo*Wh + hi*Wl in bfloat16 -- the tensor-core mode held
    to north_star's 1e-3; the program (split casts, packed [Wh|Wh][Wl] weights, memory plan) reproduces the reference."""
    case = next(c for c in golden_cases() if c["name"] == name)
    g = load_golden(name)
    m, w, cfg = build_case_model(case["backbone"], case["weight_seed"])
    B, H, W = case["B"], case["H"], case["W"]
    images, kp2d, crop = protocol.make_inputs(B, H, W, case["input_seed"])
    prog = program.build_forward_program(case["backbone"], getattr(m.backbone, "cfg", None), m._pf_cfg,
                                         {k: tuple(v.shape) for k, v in w.items()}, B, H, W, "bf16x3", use_tc=True)
    convs = [op for op in prog.ops if op.kind == lib.OP_CONV2D]
    assert sum(op.i[12] == lib.IMPL_TCGEN05 and op.i[18] == 1 for op in convs) == len(convs) - 1      # all but the 3-channel stem
    it = interp.Interp(prog, w)
    it.t(prog.inputs["images"]).copy_(images)
    it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
    it.t(prog.inputs["ref"]).copy_(torch.from_numpy(g["crop_after"]).reshape(-1, 2))
    it.run()
    assert rel_l2(it.t(prog.outputs["out"]).view(B, 1, 17, 3), g["out"]) < 1e-4


def test_memory_plan_reuses_and_never_aliases_live_buffers():
    m, w, _ = build_case_model("hrnet_32", 0)
    prog = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, {k: tuple(v.shape) for k, v in w.items()},
                                         2, 64, 64, "fp32")
    assign, slots = program.plan_memory(prog)
    roots = {b.root for op in prog.ops for b in list(op.ins) + list(op.outs) if isinstance(b, program.Buf)}
    assert sum(slots) < 0.25 * sum(r.nbytes for r in roots)          # reuse happens
    # liveness check: two roots sharing a slot must have disjoint [first def, last use] ranges
    first, last = {}, {}
    for k, op in enumerate(prog.ops):
        for b in list(op.ins) + list(op.outs):
            if isinstance(b, program.Buf):
                first.setdefault(b.root, k)
                last[b.root] = k
    by_slot = {}
    for r, s in assign.items():
        by_slot.setdefault(s, []).append((first[r], last[r]))
    residual_of, in_place = {}, 0
    for k, op in enumerate(prog.ops):
        if op.kind == lib.OP_CONV2D and isinstance(op.ins[3], program.Buf):
            r = op.ins[3].root
            residual_of[k] = [(first[r], last[r])]
            in_place += assign[r] == assign[op.outs[0].root] and r is not op.outs[0].root
    assert in_place >= 20                                            # dying residuals hand their slot to the output
    for s, iv in by_slot.items():
        iv.sort()
        for (a0, a1), (b0, b1) in zip(iv, iv[1:]):
            # the one permitted touch: op b0 reads the earlier tenant as its residual and overwrites it in place
            handover = a1 == b0 and (a0, a1) in residual_of.get(b0, ())
            assert a1 < b0 or handover, f"slot {s}: live ranges {a0, a1} and {b0, b1} overlap"


def test_flop_accounting_matches_survey():
    m, w, _ = build_case_model("hrnet_32", 0)
    prog = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, {k: tuple(v.shape) for k, v in w.items()},
                                         1, 256, 256, "fp32")
    assert abs(prog.flops() / 1e9 - 20.995) < 0.05        # SURVEY.md section 6 / BASELINE.md section 2
    qkv = sum(o.flops for o in prog.ops if "joint_blocks" in o.tag and o.tag.endswith("attn.qkv"))
    assert abs(qkv / 1e9 - 0.1671) < 1e-3


def test_lane_parallel_memory_plan_is_race_free():
    """HRNet branches run on parallel lanes (program.assign_lanes).  Two buffers may share a storage slot only if every
    op touching the earlier tenant is ordered (by the event schedule's vector clocks) before the later tenant's
    producer; and every cross-lane data dependency is covered by a wait."""
    import capf_b200
    cfg = capf_b200.make_config("hrnet_32")
    model = capf_b200.CA_PF(cfg, precision="fp16")
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    prog = program.build_forward_program("hrnet_32", model.backbone.cfg, model._pf_cfg, shapes, 4, 128, 96, "fp16", use_tc=True)
    lanes = {op.lane for op in prog.ops}
    assert lanes == {0, 1, 2, 3}
    vc, waits = program.op_clocks(prog)
    assign, _ = program.plan_memory(prog)
    touch = {}
    for k, op in enumerate(prog.ops):
        for b in list(op.ins) + list(op.outs):
            if isinstance(b, program.Buf):
                touch.setdefault(b.root, []).append(k)
    by_slot = {}
    for root, slot in assign.items():
        by_slot.setdefault(slot, []).append(root)
    shared = 0
    for slot, roots in by_slot.items():
        roots.sort(key=lambda r: touch[r][0])
        for a, b in zip(roots, roots[1:]):
            first_b = touch[b][0]
            for u in touch[a]:
                if u == first_b:       # in-place residual: the op that reads tenant a is the one that writes tenant b
                    assert prog.ops[u].ins[3].root is a and touch[a][-1] == u
                    continue
                assert vc[first_b][prog.ops[u].lane] >= u, (a.name, b.name, u, first_b)
            shared += 1
    assert shared > 50
    # read-after-write / write-after-write / write-after-read across lanes are always ordered, byte range by byte range (the
    # per-level Linears of the lifter write disjoint slabs of one token buffer from four lanes)
    def span(b):
        lo = b.root_offset * program._ITEMSIZE[b.dtype]
        return b.root, lo, lo + b.nbytes

    writes, reads = [], []
    for k, op in enumerate(prog.ops):
        ins = [span(b) for b in op.ins if isinstance(b, program.Buf)]
        outs = [span(b) for b in op.outs if isinstance(b, program.Buf)]
        for r, lo, hi in ins:
            for w, wr, a, z in writes:
                if wr is r and a < hi and lo < z:
                    assert vc[k][prog.ops[w].lane] >= w, (k, w)
        for r, lo, hi in outs:
            for w, wr, a, z in writes + reads:
                if wr is r and a < hi and lo < z and w != k:
                    assert vc[k][prog.ops[w].lane] >= w, (k, w)
        reads += [(k, r, lo, hi) for r, lo, hi in ins]
        writes += [(k, r, lo, hi) for r, lo, hi in outs]
    assert any(op.lane == 3 and "embed_proj.3" in op.tag for op in prog.ops)


def test_state_fingerprint_and_plan_cache_hygiene():
    """the weight-change detector must see tensor replacement and in-place loads (a sum of version counters
    cancels / misses them); plans are a cache -- never deep-copied or pickled; invalidate_weights() forces a repack."""
    import pickle
    import capf_b200
    from capf_b200.mvn.models._runtime import state_version
    m = capf_b200.CA_PF(capf_b200.make_config("hrnet_32"))
    v0 = state_version(m)
    assert state_version(m) == v0
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict(sd)                                   # in-place copy_: version counters move
    v1 = state_version(m)
    assert v1 != v0
    m.load_state_dict(sd, assign=True)                      # tensors replaced: ids / storage move
    v2 = state_version(m)
    assert v2 != v1
    p, q = list(m.volume_net.parameters())[:2]
    with torch.no_grad():
        p.add_(1.0)
    assert state_version(m) != v2
    # a fake cached plan (ctypes handles cannot be pickled / deep-copied)
    import ctypes
    m._plans[("fake",)] = [ctypes.c_void_p(1), state_version(m)]
    m.backbone.__dict__.setdefault("_plans", {})[("fake",)] = [ctypes.c_void_p(2), 0]
    c = copy.deepcopy(m)
    assert c._plans == {} and c.backbone.__dict__.get("_plans", {}) == {} and len(m._plans) == 1
    assert torch.equal(c.volume_net.head[1].weight, m.volume_net.head[1].weight)
    r = pickle.loads(pickle.dumps(m))
    assert r._plans == {} and r.backbone.__dict__.get("_plans", {}) == {}
    m.invalidate_weights()
    assert m._plans[("fake",)][1] is None and m.backbone._plans[("fake",)][1] is None
    m.to("cpu")
    assert m._plans == {}


def test_forward_rejects_misshaped_crop_without_gpu():
    import capf_b200
    m = capf_b200.CA_PF(capf_b200.make_config("hrnet_32")).eval()
    with pytest.raises(lib.CapfError):           # CPU tensors: there is no CPU path
        m(torch.zeros(1, 64, 64, 3), torch.zeros(1, 17, 2), torch.zeros(1, 17, 2))


def test_sibling_convs_become_one_gemm_with_output_segments(monkeypatch):
    """program.fuse_siblings: the fuse-layer convolutions of a HighResolutionModule that start from the same branch
    (pose_hrnet.py:235-277) run as one GEMM over the Cout-concatenated weights, one dense tensor per sibling.  The program with
    the merged ops computes exactly what the separate convolutions compute (TEST interpreter, same memory plan rules)."""
    case = next(c for c in golden_cases() if c["name"] == "hrnet32_b2_128x96")
    g = load_golden(case["name"])
    m, w, cfg = build_case_model(case["backbone"], case["weight_seed"])
    B, H, W = case["B"], case["H"], case["W"]
    images, kp2d, crop = protocol.make_inputs(B, H, W, case["input_seed"])
    shapes = {k: tuple(v.shape) for k, v in w.items()}
    outs, progs = [], []
    for flag in ("1", "0"):
        monkeypatch.setenv("CAPF_FUSE_SIBLINGS", flag)
        prog = program.build_forward_program("hrnet_32", m.backbone.cfg, m._pf_cfg, shapes, B, H, W, "fp16", use_tc=True)
        it = interp.Interp(prog, w)
        it.t(prog.inputs["images"]).copy_(images)
        it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
        it.t(prog.inputs["ref"]).copy_(torch.from_numpy(g["crop_after"]).reshape(-1, 2))
        it.run()
        outs.append(it.t(prog.outputs["out"]).clone())
        progs.append(prog)
    merged = [op for op in progs[0].ops if op.kind == lib.OP_CONV2D and len(op.i) > 20 and op.i[20] > 1]
    # stage3: 4 modules x {branch 0: 2 stride-2 convs, branch 2: 2 1x1 convs}; stage4: 2 multi-output modules x {branch 0: 3,
    # branch 1: 2, branch 2: 2, branch 3: 3}
    assert sorted(op.i[20] for op in merged) == [2] * 12 + [3] * 4
    assert len(progs[1].ops) - len(progs[0].ops) == sum(op.i[20] - 1 for op in merged) == 20
    for op in merged:
        widths = [op.i[21 + s] for s in range(op.i[20] - 1)]
        widths.append(op.i[4] - sum(widths))
        assert [o.shape[-1] for o in op.outs] == widths and op.i[4] <= program.SIBLING_MAX_COUT
        assert op.lane == int(re.search(r"fuse_layers\.\d+\.(\d+)\.", op.tag).group(1))
    assert sum(op.flops for op in progs[0].ops) == sum(op.flops for op in progs[1].ops)
    assert torch.equal(outs[0], outs[1])
    assert rel_l2(outs[0].view(B, 1, 17, 3), g["out"]) < 2.5e-3
