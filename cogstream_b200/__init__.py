"""cogstream_b200 -- B200-native (sm_100a) Visual Stream Compression for CogReasoner.

Drop-in surface (same names / signatures as the reference's model/kmeans_with_time.py and the
compression functions of model/cogreasoner_chat.py):

    from cogstream_b200 import kmeans_with_time_min_max, select_additional_frames
    from cogstream_b200 import VisualStreamCompressionMixin, install

plus the fused ``StreamCompressor`` pipeline and the raw op layer (``cogstream_b200.ops``) over the
C ABI in ``include/cogstream_b200.h``.  Importing the package needs neither a GPU nor the built
library; calling any op does (there is no CPU fallback).
"""
from .compression import (VisualStreamCompressionMixin, create_visual_summary_prompt, install,  # noqa: F401
                          select_additional_frames)
from . import custom_ops  # noqa: F401  (registers torch.ops.cogstream_b200.*)
from .kmeans_with_time import kmeans_with_time_min_max  # noqa: F401
from .cache import StreamCache  # noqa: F401
from .serving import StreamInputs, StreamPool, check_rows, run_sharded  # noqa: F401
from .summary_forward import make_varlen_summary_forward  # noqa: F401
from .encoder_tail import FusedEncoderTailMixin, encoder_tail, install_encoder_tail  # noqa: F401
from .sharded import FrameShardedKMeans, frame_range  # noqa: F401
from .pipeline import STATS_FIELDS, StreamCompressor, StreamConfig, gather_stats, pad_rows, shard_streams  # noqa: F401

__version__ = "0.1.0"
