"""Dataset lookup with the reference's rule (datasets/__init__.py:18-20): alias -> module -> `Dataset`.
Only a synthetic generator of the pair-file format ships here (file I/O is out of scope, SURVEY.md §2
row 10); the reference's own `davis_sequence` / `shutterstock` datasets plug in unchanged because the
Model consumes the same batch dict."""
import importlib


def get_dataset(alias):
    return importlib.import_module(__name__ + '.' + (alias or 'synthetic_sequence')).Dataset
