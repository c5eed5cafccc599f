from .converters import TFMLabelConverter  # noqa: F401  (doc2tex/modules/converter/tfm_converter.py)
