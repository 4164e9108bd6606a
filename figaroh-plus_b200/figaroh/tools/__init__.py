from . import qrdecomposition, regressor, robot  # noqa: F401
