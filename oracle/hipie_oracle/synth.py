"""Seeded synthetic weights / inputs for parity tests and the CPU baseline (test infrastructure).

No checkpoints, tokenizer vocab or datasets are available offline (SURVEY.md §7 "hard parts"), so:
  * weights follow the reference initialisers (done in the module constructors) and every tensor the
    reference zero-initialises (sampling_offsets.weight, attention_weights.*, last bbox_embed layers,
    rel_pos_*, pos_embed ...) gets N(0, 0.02) noise so that no stage is input-independent;
  * token ids are synthesised: [101] + per-class 1-3 random word-piece ids joined by 1012 + [102], zero padded.
"""
import torch


def perturb_(model, seed=0, std=0.02):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith("text_encoder"):
                continue
            zero_like = p.numel() > 1 and float(p.abs().max()) == 0.0
            if zero_like or name.endswith(("rel_pos_h", "rel_pos_w", "pos_embed", "sampling_offsets.weight",
                                           "attention_weights.weight", "attention_weights.bias")):
                p.add_(torch.randn(p.shape, generator=g) * std)
            if name.endswith("bias") and p.dim() == 1 and float(p.abs().max()) == 0.0:
                p.add_(torch.randn(p.shape, generator=g) * std)
        # FrozenBN statistics of the R50 path: make them non-trivial
        for name, b in model.named_buffers():
            if name.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.05)
            elif name.endswith("running_var"):
                b.copy_(1.0 + torch.rand(b.shape, generator=g) * 0.2)
    return model


def make_text(num_classes, max_len, seed=0, thing_fraction=0.6):
    """-> input_ids (1, L), attention_mask (1, L), positive_map {1-based class: [token idx]}, is_thing {cls: bool}."""
    g = torch.Generator().manual_seed(seed)
    ids, pos_map = [101], {}
    for c in range(1, num_classes + 1):
        n = int(torch.randint(1, 4, (1,), generator=g))
        if len(ids) + n + 2 > max_len:
            n = 1
        toks = torch.randint(1996, 30000, (n,), generator=g).tolist()
        pos_map[c] = list(range(len(ids), len(ids) + n))
        ids += toks + [1012]
    ids.append(102)
    assert len(ids) <= max_len, f"prompt of {len(ids)} tokens exceeds MAX_QUERY_LEN {max_len}"
    L = len(ids)
    input_ids = torch.zeros(1, max_len, dtype=torch.long)
    input_ids[0, :L] = torch.tensor(ids)
    attn = torch.zeros(1, max_len, dtype=torch.long)
    attn[0, :L] = 1
    n_thing = -(-num_classes * 6 // 10) if thing_fraction == 0.6 else int(num_classes * thing_fraction)
    is_thing = {c: c <= n_thing for c in range(1, num_classes + 1)}
    return input_ids, attn, pos_map, is_thing


def make_images(batch, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(3, h, w, generator=g) * 255.0 for _ in range(batch)]


def make_batch(batch, h, w, num_classes, max_len, task="detection", seed=0):
    imgs = make_images(batch, h, w, seed)
    input_ids, attn, pos_map, is_thing = make_text(num_classes if task == "detection" else 1, max_len, seed)
    if task == "grounding":
        is_thing = {1: True}
    inputs = [dict(image=im, height=h, width=w, task=task, is_thing=is_thing, positive_map_label_to_token=pos_map) for im in imgs]
    return inputs, input_ids.repeat(batch, 1), attn.repeat(batch, 1)


def fill_by_name_(module, seed=0):
    """Deterministic parameters keyed by parameter NAME and shape (used to give a reference module and the oracle's restatement
    of it identical weights without storing a state_dict in the fixture: the fill only agrees if names and shapes agree, so it
    also checks state_dict compatibility).  matrices ~ N(0, 1/fan_in), norm weights ~ 1 + 0.1 N, everything else ~ 0.1 N."""
    import zlib
    with torch.no_grad():
        items = list(module.named_parameters()) + [(n, b) for n, b in module.named_buffers() if b.dtype.is_floating_point]
        for name, p in sorted(items, key=lambda kv: kv[0]):
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
            r = torch.randn(p.shape, generator=g)
            if "bg_query_refs" in name:
                p.copy_(0.2 + 0.6 * torch.rand(p.shape, generator=g))
            elif name.endswith("running_var"):
                p.copy_(1.0 + 0.2 * torch.rand(p.shape, generator=g))
            elif p.dim() > 1:
                p.copy_(r / (p.shape[1:].numel() ** 0.5))
            elif name.endswith("weight") and p.numel() > 1:
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(0.1 * r)
    return module
