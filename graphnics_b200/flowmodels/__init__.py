from .flow_models import *      # noqa: F401,F403
from .parameters import *       # noqa: F401,F403
from .time_stepping import *    # noqa: F401,F403
