"""Special-token table the model and its callers share (values as in the reference's groma/constants.py:5-25)."""
# serve-layer constants (groma/constants.py:1-3)
CONTROLLER_HEART_BEAT_EXPIRATION = 30
WORKER_HEART_BEAT_INTERVAL = 15
LOGDIR = "."

IGNORE_INDEX = -100

_NAMES = ["pad", "bos", "eos", "unk", "sep", "boi", "eoi", "bor", "eor", "boe", "eoe", "image", "region", "rbox", "gbox",
          "rfeat", "ground"]
_TOKENS = ["[PAD]", "<s>", "</s>", "<unk>", "<sep>", "<img>", "</img>", "<roi>", "</roi>", "<p>", "</p>", "<image>",
           "<region>", "<refer_box>", "<ground_box>", "<refer_feat>", "[grounding]"]
DEFAULT_TOKENS = dict(zip(_NAMES, _TOKENS))
REGION_IDX_TOKENS = [f"<r{i}>" for i in range(100)]
