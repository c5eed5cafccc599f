from .converters import AttnLabelConverter  # noqa: F401  (doc2tex/modules/converter/attn_converter.py)
