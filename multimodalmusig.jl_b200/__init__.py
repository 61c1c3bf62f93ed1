"""mmsig-b200: B200-native variational-EM inner loop of MMCTM / CTM / LDA
(drop-in for `fit!` of shahcompbio/MultiModalMuSig.jl)."""
from . import counts, synth  # noqa: F401
