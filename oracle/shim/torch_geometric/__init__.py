"""Oracle shim: minimal stand-in for the (absent, unpinned) torch_geometric ~1.1-1.3 API
the reference imports.  TEST INFRASTRUCTURE ONLY -- never imported by the product."""
from . import data, utils  # noqa: F401
