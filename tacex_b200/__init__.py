"""tacex_b200 -- B200-native tactile-image synthesis engine (GelSight Mini: Taxim RGB + FOTS markers + gel FEM).

Drop-in for the tactile hot path of DH-Ng/TacEx: hand-written sm_100a CUDA kernels behind a C ABI
(include/tacex_b200.h), bound by subclasses of the reference's GelSightSimulator plug-in interface.
"""

__version__ = "0.1.0"
