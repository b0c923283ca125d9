"""Mirror of the reference's ``precompute`` package (``from precompute import propagation``,
/root/reference/model.py:9)."""
from . import propagation  # noqa: F401
