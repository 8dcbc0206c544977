"""Plug-in lookup with the reference's rule (models/__init__.py:18-20): alias -> module -> `Model`."""
import importlib


def get_model(alias):
    return importlib.import_module(__name__ + '.' + alias).Model
