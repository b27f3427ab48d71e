"""reformer_tts_b200: B200-native (sm_100a) implementation of Reformer-TTS's LSH-attention / reversible /
chunked-FFN hot path behind the reference's module API.  See DESIGN.md."""
__version__ = "0.1.0"
