"""Prompt / positive-map builder of the HIPIE inference edge.

Mirrors /root/reference/projects/HIPIE/hipie/data/coco_dataset_mapper_uni.py:54-91 (`create_queries_and_maps`), :732-736
(`clean_name`), :1024-1058 (`create_positive_dict`) and hipie/data/datasets/catids.py:3-43 (`get_openseg_labels`): the category names
are joined with ". " into one caption, tokenised with the BERT word-piece tokenizer, and every class is mapped to the token
positions its characters cover; `HIPIE_IMG.inference` pools the grounding logits over exactly those positions.

The tokenizer is HuggingFace's BertTokenizerFast built from the `vocab.txt` of the `bert-base-uncased` directory the reference
expects under projects/HIPIE/ (assets/INSTALL.md); no network access is needed, only that file.
"""
import os
import re
from collections import defaultdict

DEFAULT_BERT_DIRS = ("projects/HIPIE/bert-base-uncased", os.environ.get("HIPIE_BERT_DIR", ""))


def load_tokenizer(path=None):
    """BertTokenizerFast from a local bert-base-uncased directory (or a vocab.txt path).  Raises with the searched locations
    when the vocabulary is not there -- there is no built-in vocabulary."""
    from transformers import BertTokenizerFast
    cands = [path] if path else [p for p in DEFAULT_BERT_DIRS if p]
    for c in cands:
        vocab = c if c.endswith(".txt") else os.path.join(c, "vocab.txt")
        if os.path.isfile(vocab):
            with open(vocab, "r", encoding="utf-8") as f:
                table = {tok.rstrip("\n"): i for i, tok in enumerate(f)}
            try:                                   # transformers >= 5: the vocabulary is passed as a dict
                tok = BertTokenizerFast(vocab=table, do_lower_case=True)
                if tok.vocab_size == len(table):
                    return tok
            except TypeError:
                pass
            return BertTokenizerFast(vocab_file=vocab, do_lower_case=True)      # transformers 4.x
    raise FileNotFoundError(f"BERT vocabulary not found (looked for vocab.txt in {cands}); put bert-base-uncased under projects/HIPIE/ "
                            "as the reference's INSTALL.md says, or set HIPIE_BERT_DIR")


def clean_name(name):
    name = re.sub(r"\(.*\)", "", name)
    name = re.sub(r"_", " ", name)
    name = re.sub(r"  ", " ", name)
    return name


def create_positive_dict(tokenized, tokens_positive, labels):
    """positive_map[token position] = label and label -> [token positions], with the reference's +-1/2/3 character fallbacks
    when a span boundary falls on a character that belongs to no token (spaces)."""
    positive_map = defaultdict(int)
    positive_map_label_to_token = {}
    for j, tok_list in enumerate(tokens_positive):
        for (beg, end) in tok_list:
            beg_pos = tokenized.char_to_token(beg)
            end_pos = tokenized.char_to_token(end - 1)
            if beg_pos is None:
                try:
                    beg_pos = tokenized.char_to_token(beg + 1)
                    if beg_pos is None:
                        beg_pos = tokenized.char_to_token(beg + 2)
                except Exception:
                    beg_pos = None
            if end_pos is None:
                try:
                    end_pos = tokenized.char_to_token(end - 2)
                    if end_pos is None:
                        end_pos = tokenized.char_to_token(end - 3)
                except Exception:
                    end_pos = None
            if beg_pos is None or end_pos is None:
                continue
            positive_map_label_to_token[labels[j]] = []
            for i in range(beg_pos, end_pos + 1):
                positive_map[i] = labels[j]
                positive_map_label_to_token[labels[j]].append(i)
    return positive_map, positive_map_label_to_token


def create_queries_and_maps(categories, tokenizer, separation_tokens=". ", things_only=False):
    """categories: [{"name": str, optional "isthing": 0/1}, ...] -> (caption, {1-based label: [token positions]})"""
    label_list = []
    for x in categories:
        isthing = x["isthing"] if "isthing" in x else 1
        if isthing or (not things_only):
            label_list.append(x["name"])
    labels = list(range(1, len(label_list) + 1))
    label_list = [clean_name(i) for i in label_list]
    tokens_positive = []
    objects_query = ""
    separation_tokens = ". "          # the reference overrides the argument: training always used ". "
    for _index, label in enumerate(label_list):
        start_i = len(objects_query)
        objects_query += label
        end_i = len(objects_query)
        tokens_positive.append([(start_i, end_i)])
        if _index != len(label_list) - 1:
            objects_query += separation_tokens
    tokenized = tokenizer(objects_query, return_tensors="pt")
    _, positive_map_label_to_token = create_positive_dict(tokenized, tokens_positive, labels=labels)
    return objects_query, positive_map_label_to_token


def tokenize_captions(tokenizer, captions, max_query_len, pad_max=True):
    """HIPIE_IMG.forward_text's tokenisation (hipie_img.py:903-912): -> (input_ids, attention_mask, sep token id of '.')"""
    # the reference calls tokenizer.batch_encode_plus; __call__ is the same entry point in every transformers release
    tokenized = tokenizer(list(captions), max_length=max_query_len, padding="max_length" if pad_max else "longest",
                          return_special_tokens_mask=True, return_tensors="pt", truncation=True)
    sep_token = tokenizer(".").input_ids[1]
    return tokenized.input_ids, tokenized.attention_mask, sep_token


def get_openseg_labels(dataset, prompt_engineered=False, root=None):
    """Category list of an open-vocabulary benchmark from the reference's `id:name` label files
    (hipie/data/datasets/openseg_labels/<dataset>[_with_prompt_eng].txt).  The files are data of the reference checkout and
    are read from there (`root`, $HIPIE_LABELS_DIR, or projects/HIPIE/hipie/data/datasets/openseg_labels)."""
    roots = [root, os.environ.get("HIPIE_LABELS_DIR"), "projects/HIPIE/hipie/data/datasets/openseg_labels"]
    fname = f"{dataset}_with_prompt_eng.txt" if prompt_engineered else f"{dataset}.txt"
    for r in roots:
        if r and os.path.isfile(os.path.join(r, fname)):
            with open(os.path.join(r, fname), "r") as f:
                lines = f.read().splitlines()
            categories = []
            for line in lines:
                id_, name = line.split(":", maxsplit=1)
                if name == "invalid_class_id":
                    continue
                categories.append({"id": int(id_), "name": name})
            return categories
    raise FileNotFoundError(f"{fname} not found under {[r for r in roots if r]}")
