from .converters import AttnLabelConverter, TFMLabelConverter, create_converter  # noqa: F401
