from .narde import Narde, rotate_board  # noqa: F401
from .narde_env import NardeEnv  # noqa: F401
