"""Model registry surface of the reference (`models.utils`, `models.ncsnpp`, `models.ema`)."""
from . import utils  # noqa: F401
from . import ncsnpp  # noqa: F401  (registers 'ncsnpp')
from . import ema  # noqa: F401
