"""Host-side mirror of ``reformer_tts.model`` for the hot path (same class names, constructor kwargs, forward
signatures and state-dict keys as ref:reformer_tts/model/{reformer,reversible,modules,reformer_tts}.py)."""
from .modules import FeedForward, EncoderPreNet, DecoderPreNet, PostConvNet, ScaledPositionalEncoding  # noqa: F401
from .reformer import (Chunk, LSHSelfAttentionWrapper, MultiheadAttentionWrapper, ReformerDec, ReformerEnc,  # noqa: F401
                       WithNorm)
from .reformer_tts import ReformerTTS, pad_to_multiple  # noqa: F401
from .reversible import (Deterministic, ReversibleBlock, ReversibleHalfResidual, ReversibleSequence,  # noqa: F401
                         ReversibleSwap)
