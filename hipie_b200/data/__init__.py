from .prompts import (clean_name, create_positive_dict, create_queries_and_maps, get_openseg_labels, load_tokenizer,  # noqa: F401
                      tokenize_captions)
from .transforms import ResizeShortestEdge  # noqa: F401
