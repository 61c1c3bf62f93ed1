"""mmsig-b200: B200-native variational-EM inner loop of MMCTM / CTM / LDA
(drop-in for `fit!` of shahcompbio/MultiModalMuSig.jl)."""
from . import capi, counts, io, restarts, synth  # noqa: F401
from .counts import format_counts_ctm, format_counts_lda, format_counts_mmctm  # noqa: F401
from .models import ILDA, IMMCTM, LDA, MMCTM, MMCTMGroup  # noqa: F401
