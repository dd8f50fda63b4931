"""Generate golden fixtures under tests/golden/ from the REAL reference code (run in the build container).

Only the leaf modules of /root/reference that import with `sys.modules` stubs are used:
  * ops/functions/ms_deform_attn_func.py  -> ms_deform_attn_core_pytorch   (pins row a12)
  * backbone/utils.py                     -> window_partition/unpartition, get_rel_pos,
                                             add_decomposed_rel_pos, get_abs_pos (pins parts of a2/a3)
  * deformable_detr/position_encoding.py, maskdino/pixel_decoder/position_encoding.py (a5/a17)
  * maskdino/utils/utils.py               -> gen_encoder_output_proposals, gen_sineembed_for_position
  * deformable_detr/fuse_helper.py        -> BiMultiHeadAttention (a9)
The GPU box has no /root/reference, hence the committed fixtures.
Usage: python oracle/gen_golden.py
"""
import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference/projects/HIPIE/hipie"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load(name, path, package_stubs=()):
    for stub in package_stubs:
        if stub not in sys.modules:
            m = types.ModuleType(stub)
            m.__path__ = []
            sys.modules[stub] = m
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- MSDeformAttn core (reference ops/test.py fixture: seed 3, N,M,D=1,2,2 Lq,L,P=2,2,2 shapes (6,4),(3,2))
    sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
    f = _load("ref_msda_func", f"{REF}/models/deformable_detr/ops/functions/ms_deform_attn_func.py")
    core = f.ms_deform_attn_core_pytorch
    cases = {}
    torch.manual_seed(3)
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    for tag, dt in (("test_py_double", torch.float64), ("test_py_float", torch.float32)):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        w = torch.rand(N, Lq, M, L, P) + 1e-5
        w /= w.sum(-1, keepdim=True).sum(-2, keepdim=True)
        out = core(value.to(dt), shapes, loc.to(dt), w.to(dt))
        cases[tag] = dict(value=value.to(dt), shapes=shapes, loc=loc.to(dt), w=w.to(dt), out=out)
    # HIPIE-shaped case (D=32, L=4, P=4, M=8) incl. out-of-range locations
    g = torch.Generator().manual_seed(11)
    shapes2 = torch.as_tensor([(16, 12), (8, 6), (4, 3), (2, 2)], dtype=torch.long)
    S2 = int(shapes2.prod(1).sum())
    value = torch.randn(2, S2, 8, 32, generator=g)
    loc = torch.rand(2, 37, 8, 4, 4, 2, generator=g) * 1.3 - 0.15
    w = torch.softmax(torch.randn(2, 37, 8, 16, generator=g), -1).view(2, 37, 8, 4, 4)
    cases["hipie_shape_f32"] = dict(value=value, shapes=shapes2, loc=loc, w=w, out=core(value, shapes2, loc, w))
    torch.save(cases, os.path.join(OUT, "msda_core.pt"))

    # --- ViT helpers
    u = _load("ref_vit_utils", f"{REF}/backbone/utils.py")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 20, 24, 8, generator=g)
    win, pad_hw = u.window_partition(x, 14)
    back = u.window_unpartition(win, 14, pad_hw, (20, 24))
    q = torch.randn(3, 6 * 5, 16, generator=g)
    attn = torch.randn(3, 30, 30, generator=g)
    rph = torch.randn(2 * 7 - 1, 16, generator=g)   # needs interpolation (13 != 11)
    rpw = torch.randn(2 * 5 - 1, 16, generator=g)
    attn_out = u.add_decomposed_rel_pos(attn.clone(), q, rph, rpw, (6, 5), (6, 5))
    abs_pos = torch.randn(1, 14 * 14 + 1, 8, generator=g)
    ap = u.get_abs_pos(abs_pos, True, (20, 24))
    torch.save(dict(x=x, win=win, pad_hw=pad_hw, back=back, q=q, attn=attn, rph=rph, rpw=rpw, attn_out=attn_out,
                    Rh=u.get_rel_pos(6, 6, rph), Rw=u.get_rel_pos(5, 5, rpw), abs_pos=abs_pos, abs_pos_out=ap),
               os.path.join(OUT, "vit_utils.pt"))

    # --- position encodings (two variants) and MaskDINO utils
    misc = types.ModuleType("hipie_util_misc")

    class NestedTensor:  # minimal stand-in for util/misc.py:288 (only .tensors/.mask are read)
        def __init__(self, tensors, mask):
            self.tensors, self.mask = tensors, mask
    misc.NestedTensor = NestedTensor
    for pkg in ("refpkg", "refpkg.models", "refpkg.models.deformable_detr", "refpkg.util"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m
    sys.modules["refpkg.util.misc"] = misc
    spec = importlib.util.spec_from_file_location("refpkg.models.deformable_detr.position_encoding",
                                                  f"{REF}/models/deformable_detr/position_encoding.py")
    pe = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = pe
    spec.loader.exec_module(pe)
    mask = torch.zeros(2, 9, 11, dtype=torch.bool)
    mask[1, 7:, :] = True
    mask[1, :, 8:] = True
    xt = torch.zeros(2, 4, 9, 11)
    pos_detr = pe.PositionEmbeddingSine(128, normalize=True)(NestedTensor(xt, mask))
    pe2 = _load("ref_md_pe", f"{REF}/models/maskdino/pixel_decoder/position_encoding.py")
    pos_md = pe2.PositionEmbeddingSine(128, normalize=True)(xt)
    mu = _load("ref_md_utils", f"{REF}/models/maskdino/utils/utils.py")
    mem = torch.randn(2, 9 * 11 + 5 * 6, 16, generator=g)
    pmask = torch.zeros(2, 9 * 11 + 5 * 6, dtype=torch.bool)
    ss = torch.as_tensor([(9, 11), (5, 6)])
    om, op = mu.gen_encoder_output_proposals(mem, pmask, ss)
    pos4 = torch.rand(7, 2, 4, generator=g)
    sine4 = mu.gen_sineembed_for_position(pos4)
    torch.save(dict(mask=mask, pos_detr=pos_detr, pos_md=pos_md, mem=mem, ss=ss, out_mem=om, out_prop=op, pos4=pos4,
                    sine4=sine4, inv_sig_in=torch.tensor([0.0, 1e-7, 0.3, 0.999999, 1.0]),
                    inv_sig_out=mu.inverse_sigmoid(torch.tensor([0.0, 1e-7, 0.3, 0.999999, 1.0]))),
               os.path.join(OUT, "posenc_utils.pt"))

    # --- BiMultiHeadAttention (VL fusion core)
    import transformers  # noqa: F401  (must be imported before stubbing timm)
    timm = types.ModuleType("timm"); timm.__path__ = []
    tm = types.ModuleType("timm.models"); tm.__path__ = []
    tl = types.ModuleType("timm.models.layers")

    class DropPath(torch.nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()
    tl.DropPath = DropPath
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    fh = _load("ref_fuse_helper", f"{REF}/models/deformable_detr/fuse_helper.py")

    class _C:
        pass
    cfg = _C(); cfg.MODEL = _C(); cfg.MODEL.DYHEAD = _C(); cfg.MODEL.DYHEAD.FUSE_CONFIG = _C()
    cfg.MODEL.DYHEAD.FUSE_CONFIG.STABLE_SOFTMAX_2D = False
    cfg.MODEL.DYHEAD.FUSE_CONFIG.CLAMP_MIN_FOR_UNDERFLOW = True
    cfg.MODEL.DYHEAD.FUSE_CONFIG.CLAMP_MAX_FOR_OVERFLOW = True
    torch.manual_seed(7)
    blk = fh.BiAttentionBlockForCheckpoint(v_dim=32, l_dim=48, embed_dim=64, num_heads=4, dropout=0.1, drop_path=0.0,
                                           init_values=1.0 / 6, cfg=cfg).eval()
    v = torch.randn(2, 50, 32)
    l = torch.randn(2, 9, 48)
    am = torch.ones(2, 9, dtype=torch.long)
    am[1, 6:] = 0
    with torch.no_grad():
        ov, ol = blk(v, l, attention_mask_l=am)
    torch.save(dict(state=blk.state_dict(), v=v, l=l, mask=am, out_v=ov, out_l=ol), os.path.join(OUT, "vlfuse_block.pt"))
    print("golden fixtures written to", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
