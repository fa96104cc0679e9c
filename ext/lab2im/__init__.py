from . import utils  # noqa: F401
from . import edit_volumes  # noqa: F401
