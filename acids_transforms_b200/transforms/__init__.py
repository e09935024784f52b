from .base import *        # noqa: F401,F403
from .base import AudioTransform, ComposeAudioTransform, NotInvertibleError, InversionEnumType  # noqa: F401
from .raw import *         # noqa: F401,F403
from .stft import *        # noqa: F401,F403
from .dgt import *         # noqa: F401,F403
from .norm import *        # noqa: F401,F403
from .spectral_repr import *  # noqa: F401,F403
from .mel import *         # noqa: F401,F403
from .misc import *        # noqa: F401,F403
from .oadd import *        # noqa: F401,F403
