"""Re-export of the path configuration (the oracle is sized by the same dataclass as the product)."""
from groma_b200.config import PathConfig, tiny_config, SyntheticTokenizer, NEW_TOKENS  # noqa: F401
