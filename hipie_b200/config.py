"""Configuration system with the detectron2 / yacs surface the HIPIE project uses:
`get_cfg()`, `add_hipie_config(cfg)`, `cfg.merge_from_file(yaml with _BASE_)`, `cfg.merge_from_list([...])`,
attribute access, `clone()`, `freeze()` — so `projects/HIPIE/configs/**` load unchanged
(/root/reference/detectron2/config/config.py, /root/reference/projects/HIPIE/hipie/config.py:5-284,
 /root/reference/projects/HIPIE/hipie/models/maskdino/config.py:9-151).  yacs/fvcore are not required.
Only the keys the inference path reads get explicit defaults; unknown keys from a YAML are accepted as-is.
"""
import ast
import copy
import os

import yaml

BASE_KEY = "_BASE_"


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError(f"Attempted to set {name} on a frozen CfgNode")
        self[name] = value

    def freeze(self):
        object.__setattr__(self, "_frozen", True)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        object.__setattr__(self, "_frozen", False)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def clone(self):
        c = copy.deepcopy(self)
        return c

    def __deepcopy__(self, memo):
        c = CfgNode()
        for k, v in self.items():
            dict.__setitem__(c, k, copy.deepcopy(v, memo))
        return c

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k]._merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    @staticmethod
    def load_yaml_with_base(filename):
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) or {}
        if BASE_KEY in cfg:
            base = cfg.pop(BASE_KEY)
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not any(map(base.startswith, ["/", "https://", "http://"])):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = CfgNode.load_yaml_with_base(base)

            def merge(a, b):
                for k, v in a.items():
                    if isinstance(v, dict) and k in b and isinstance(b[k], dict):
                        merge(v, b[k])
                    else:
                        b[k] = v
            merge(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename, allow_unsafe=True):
        loaded = CfgNode.load_yaml_with_base(cfg_filename)
        object.__setattr__(self, "_source_file", os.path.abspath(cfg_filename))
        self._merge(_tuples(loaded))

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0
        for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    if v.lower() in ("on", "true"):
                        v = True
                    elif v.lower() in ("off", "false"):
                        v = False
            node[parts[-1]] = v


CN = CfgNode


def _tuples(d):
    """yaml has no tuples: '(1024,)' style strings are literal-evaluated like yacs does."""
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out[k] = _tuples(v)
        elif isinstance(v, str) and v.startswith("(") and v.endswith(")"):
            try:
                out[k] = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                out[k] = v
        else:
            out[k] = v
    return out


def get_cfg() -> CfgNode:
    """The subset of detectron2's defaults (config/defaults.py) that the HIPIE yamls and the inference path touch."""
    c = CN()
    c.VERSION = 2
    c.MODEL = CN(dict(DEVICE="cuda", META_ARCHITECTURE="GeneralizedRCNN", WEIGHTS="", MASK_ON=False,
                      PIXEL_MEAN=[103.530, 116.280, 123.675], PIXEL_STD=[1.0, 1.0, 1.0],
                      BACKBONE=dict(NAME="build_resnet_backbone", FREEZE_AT=2),
                      RESNETS=dict(DEPTH=50, OUT_FEATURES=["res4"], NUM_GROUPS=1, NORM="FrozenBN", WIDTH_PER_GROUP=64,
                                   STRIDE_IN_1X1=True, RES5_DILATION=1, RES2_OUT_CHANNELS=256, STEM_OUT_CHANNELS=64,
                                   DEFORM_ON_PER_STAGE=[False, False, False, False], DEFORM_MODULATED=False,
                                   DEFORM_NUM_GROUPS=1, STEM_TYPE="basic", RES5_MULTI_GRID=[1, 1, 1]),
                      SEM_SEG_HEAD=dict(NAME="SemSegFPNHead", IN_FEATURES=["p2", "p3", "p4", "p5"], IGNORE_VALUE=255,
                                        NUM_CLASSES=54, CONVS_DIM=128, COMMON_STRIDE=4, NORM="GN", LOSS_WEIGHT=1.0)))
    c.INPUT = CN(dict(MIN_SIZE_TRAIN=(800,), MAX_SIZE_TRAIN=1333, MIN_SIZE_TEST=800, MAX_SIZE_TEST=1333, FORMAT="BGR",
                      MASK_FORMAT="polygon", RANDOM_FLIP="horizontal", MIN_SIZE_TRAIN_SAMPLING="choice",
                      CROP=dict(ENABLED=False, TYPE="relative_range", SIZE=[0.9, 0.9])))
    c.DATASETS = CN(dict(TRAIN=(), TEST=(), PROPOSAL_FILES_TRAIN=(), PROPOSAL_FILES_TEST=()))
    c.DATALOADER = CN(dict(NUM_WORKERS=4, ASPECT_RATIO_GROUPING=True, SAMPLER_TRAIN="TrainingSampler", REPEAT_THRESHOLD=0.0,
                           FILTER_EMPTY_ANNOTATIONS=True))
    c.SOLVER = CN(dict(IMS_PER_BATCH=16, BASE_LR=0.001, STEPS=(30000,), MAX_ITER=40000, WARMUP_FACTOR=0.001, WARMUP_ITERS=1000,
                       WEIGHT_DECAY=0.0001, CHECKPOINT_PERIOD=5000, AMP=dict(ENABLED=False),
                       CLIP_GRADIENTS=dict(ENABLED=False, CLIP_TYPE="value", CLIP_VALUE=1.0, NORM_TYPE=2.0)))
    c.TEST = CN(dict(EVAL_PERIOD=0, DETECTIONS_PER_IMAGE=100, EXPECTED_RESULTS=[], AUG=dict(ENABLED=False)))
    c.OUTPUT_DIR = "./output"
    c.SEED = -1
    c.CUDNN_BENCHMARK = False
    return c


def add_hipie_config(cfg):
    """Keys and defaults of hipie/config.py:5-284 that the inference path reads (others arrive from the YAMLs)."""
    m = cfg.MODEL
    cfg.UNI = True
    for k, v in dict(DECOUPLE_TGT=False, STILL_TGT_FOR_BOTH=False, CLS_POOL_TYPE="average", USE_IOU_BRANCH=False, PARALLEL_DET=False,
                     OTA=False, LANG_GUIDE_DET=True, VL_FUSION_USE_CHECKPOINT=True, USE_EARLY_FUSION=True, USE_ADDITIONAL_BERT=False,
                     LANG_AS_CLASSIFIER=True, STILL_CLS_FOR_ENCODER=False, OBJECT_MASK_THRESHOLD=0.25, OVERLAP_THRESHOLD=0.8,
                     POINT_SAMPLE=False, MODE_FREE_MATCHING_INFERENCE=False, PANO_TRANSFORM_EVAL=True, PANO_TEMPERATURE=0.06,
                     PANO_TEMPERATURE_CLIP_FG=0.06, PART_MODE=False, MAX_INSTANCES=0).items():
        m[k] = v
    m.LANGUAGE_BACKBONE = CN(dict(USE_CHECKPOINT=False, TOKENIZER_TYPE="bert-base-uncased", MODEL_TYPE="bert-base-uncased", LANG_DIM=768,
                                  MAX_QUERY_LEN=256, N_LAYERS=1, UNUSED_TOKEN=106, MASK_SPECIAL=False, PAD_MAX=True))
    m.DYHEAD = CN(dict(PRIOR_PROB=0.01, LOG_SCALE=0.0, FUSE_CONFIG=dict(CLAMP_MIN_FOR_UNDERFLOW=True, CLAMP_MAX_FOR_OVERFLOW=True,
                       CLAMP_BERTATTN_MIN_FOR_UNDERFLOW=True, CLAMP_BERTATTN_MAX_FOR_OVERFLOW=True, SEPARATE_BIDIRECTIONAL=False,
                       STABLE_SOFTMAX_2D=False, CLAMP_DOT_PRODUCT=True)))
    m.DDETRS = CN(dict(NUM_CLASSES=None, USE_CHECKPOINT=False, NHEADS=8, DROPOUT=0.1, DIM_FEEDFORWARD=2048, ENC_LAYERS=6, DEC_LAYERS=6,
                       NUM_VL_LAYERS=1, VL_HIDDEN_DIM=2048, TWO_STAGE=False, TWO_STAGE_NUM_PROPOSALS=300, TWO_STAGE_NUM_BG_PROPOSALS=0,
                       MIXED_SELECTION=False, LOOK_FORWARD_TWICE=False, CTRL_LAYERS=3, USE_DINO=False, DYNAMIC_LABEL_ENC=False,
                       HIDDEN_DIM=256, NUM_OBJECT_QUERIES=300, DEC_N_POINTS=4, ENC_N_POINTS=4, NUM_FEATURE_LEVELS=4, MASK_THRES=0.5,
                       MASK_STRIDE=4, NEW_MASK_HEAD=False, USE_RAFT=False, USE_REL_COORD=True, FORCE_NO_LOC=False,
                       BG_QUERY_FROM_LANG=False, DN_NUMBER=100, DP_NUMBER=0))
    m.CLIP = CN(dict(ENABLED=False, ENABLED_TRAIN=False, NAME="ViT-L-14-336", ALPHA=0.35, BETA=0.7, FG_IOU_A=0.3, FG_IOU_B=1.7, AGG_MODE="MUL"))
    m.MASKDINO = CN(dict(ENABLED=False, SHARE_ENCODER=False, CONFIG_PATH="", PRETRAINED="", SHARE_CLS_HEAD=False, LOSS_WEIGHT=1.0,
                         FIXED_LINEAR_HEAD=False))
    m.VIT = CN(dict(NAME="ViT-Base", OUT_FEATURES=["res3", "res4", "res5"], USE_CHECKPOINT=False))
    m.RESNETS.OUT_FEATURES = ["res3", "res4", "res5"]
    cfg.SAM = CN(dict(ENABLED=False, CHECKPOINT="", TYPE="vit_h"))
    cfg.TEST.EVAL_AFTER_TRAIN = True
    cfg.TEST.USE_BG_FOR_PANO_ON = True
    cfg.TEST.BG_CLS_AGNOSTIC = False
    cfg.TEST.MAX_POOL = False
    cfg.INPUT.DATASET_MAPPER_NAME = "detr"
    cfg.INPUT.CROP_SIZE = 1024
    cfg.FIND_UNUSED_PARAMETERS = False


def add_maskdino_config(cfg):
    """hipie/models/maskdino/config.py:9-151 — keys read by the pixel decoder / decoder construction."""
    m = cfg.MODEL
    m.MaskDINO = CN(dict(LEARN_TGT=False, PANO_BOX_LOSS=False, SEMANTIC_CE_LOSS=False, DEEP_SUPERVISION=True, NHEADS=8, DROPOUT=0.1,
                         DIM_FEEDFORWARD=2048, ENC_LAYERS=0, DEC_LAYERS=6, INITIAL_PRED=True, PRE_NORM=False, HIDDEN_DIM=256,
                         NUM_OBJECT_QUERIES=100, TWO_STAGE=True, INITIALIZE_BOX_TYPE="no", DN="seg", DN_NOISE_SCALE=0.4, DN_NUM=100,
                         ENFORCE_INPUT_PROJ=False, SIZE_DIVISIBILITY=32, DYNAMIC_LABEL_ENC=True, DYNAMIC_LABEL_ENC_DROPOUT=0.1,
                         TEST=dict(SEMANTIC_ON=True, INSTANCE_ON=False, PANOPTIC_ON=False, OBJECT_MASK_THRESHOLD=0.0, OVERLAP_THRESHOLD=0.0)))
    m.SEM_SEG_HEAD._merge(dict(MASK_DIM=256, DIM_FEEDFORWARD=1024, TRANSFORMER_ENC_LAYERS=0, PIXEL_DECODER_NAME="MaskDINOEncoder",
                               DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES=["res3", "res4", "res5"], NUM_FEATURE_LEVELS=3,
                               TOTAL_NUM_FEATURE_LEVELS=4, FEATURE_ORDER="high2low"))


def load_maskdino_cfg(cfg):
    """Second, independent cfg tree built from MODEL.MASKDINO.CONFIG_PATH (hipie/models/maskdino/build.py:8-19)."""
    path = cfg.MODEL.MASKDINO.CONFIG_PATH
    if not cfg.MODEL.MASKDINO.ENABLED or not path:
        return None
    cands = [path]
    src = getattr(cfg, "_source_file", None)
    if src:
        root = src
        for _ in range(6):
            root = os.path.dirname(root)
            cands.append(os.path.join(root, path))
    for c in cands:
        if os.path.exists(c):
            md = get_cfg()
            add_maskdino_config(md)
            md.merge_from_file(c)
            return md
    return None


def setup_cfg(config_file, opts=()):
    """get_cfg + add_hipie_config + merge (the sequence of projects/HIPIE/train_net.py:251-256)."""
    cfg = get_cfg()
    add_hipie_config(cfg)
    cfg.merge_from_file(config_file)
    cfg.merge_from_list(list(opts))
    md = load_maskdino_cfg(cfg)
    object.__setattr__(cfg, "_maskdino_cfg", md)
    return cfg
