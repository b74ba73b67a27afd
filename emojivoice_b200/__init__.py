"""emojivoice_b200: B200-native (sm_100a) synthesis hot path of rosielab/emojivoice behind the reference's own
call surface.  See DESIGN.md / INTEGRATION.md."""
from .config import EMOJI_MAPPING_FEMALE, EMOJI_MAPPING_MALE, HIFIGAN_V1, VCTK, AttrDict, MatchaConfig  # noqa: F401
from .emoji_frontend import emoji_to_spk, intersperse  # noqa: F401
from .hifigan import Denoiser, Generator, to_waveform  # noqa: F401
from .matcha import MatchaTTS  # noqa: F401
from . import audio_io, batch, monotonic_align, sharding, text_cleaners, text_frontend  # noqa: F401
from .monotonic_align import maximum_path  # noqa: F401
from .batch import Lanes, lanes_for, synthesise_corpus, synthesise_file  # noqa: F401

__version__ = "0.1.0"
