"""class_query_vad_b200 -- B200-native (sm_100a) class-query decoder hot path of dlrudco/class-query-vad.

Host side = thin Python/PyTorch mirror of the reference's operator interfaces (same class names, forward signatures,
parameter names); all arithmetic runs in hand-written CUDA behind the C ABI of include/cqvad.h (libcqvad.so).
"""
from . import _lib
from .engine import DecoderEngine, pack_decoder_weights
from .functions.decoder_func import DecoderFunction
from .functions.ms_deform_attn_func import MSDeformAttnFunction, ms_deform_attn_indices
from .modules.ms_deform_attn import MSDeformAttn3D
from .modules.attention import MultiheadAttention
from .modules.position_encoding import PositionEmbeddingSine_3D, build_position_encoding, gen_sineembed_for_position
from .modules.encoder import (DeformableTransformerEncoderLayer, DeformableTransformerEncoder, encoder_layer_forward,
                              pack_encoder_layer_weights, encoder_to_decoder_memory)
from .modules.transformer import Transformer, flatten_levels, input_proj_levels
from .modules.criterion import (HungarianMatcherAVA, SetCriterionAVA, PostProcessAVA, PostProcessUCF, PostProcessJHMDB,
                                pack_targets)
from .modules.heads import DETRHeads, HeadsFunction
from .optim import FlatAdamW
from . import detections
from .modules.neck import SimpleFeaturePyramid
from .modules.decoder import (MLP, ConvBlock, TransformerDecoderLayer, TransformerClassDecoderLayer, TransformerDecoder,
                              build_decoder)

__all__ = ["DecoderEngine", "DecoderFunction", "pack_decoder_weights", "MSDeformAttnFunction", "ms_deform_attn_indices", "MSDeformAttn3D",
           "MultiheadAttention", "PositionEmbeddingSine_3D", "build_position_encoding", "gen_sineembed_for_position",
           "MLP", "ConvBlock", "TransformerDecoderLayer", "TransformerClassDecoderLayer", "TransformerDecoder",
           "build_decoder", "DeformableTransformerEncoderLayer", "DeformableTransformerEncoder", "encoder_layer_forward",
           "pack_encoder_layer_weights", "encoder_to_decoder_memory", "Transformer", "flatten_levels", "input_proj_levels",
           "SimpleFeaturePyramid", "DETRHeads", "HeadsFunction", "FlatAdamW", "HungarianMatcherAVA", "SetCriterionAVA", "PostProcessAVA", "PostProcessUCF", "PostProcessJHMDB", "pack_targets"]
