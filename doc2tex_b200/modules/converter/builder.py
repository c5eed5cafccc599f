from .converters import create_converter  # noqa: F401  (doc2tex/modules/converter/builder.py)
