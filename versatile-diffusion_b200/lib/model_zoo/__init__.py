from .common.get_model import get_model  # noqa: F401  (same import the reference's app.py uses, app.py:28)
