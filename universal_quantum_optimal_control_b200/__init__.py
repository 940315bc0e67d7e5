"""uqoc-b200: B200-native (sm_100a) disorder-sampled unitary propagation + fidelity loss.

Drop-in for the hot path of shiminki/universal_quantum_optimal_control
(``unitary_generator`` / ``fidelity_fn`` / ``loss_fn`` / ``error_sampler`` of
``model/universal_model_trainer.py:27-33``) over the C ABI in ``include/uqoc.h``.
"""
from .ops import (  # noqa: F401
    FusedStep,
    autotune_flags,
    batched_unitary_generator,
    custom_loss,
    fidelity,
    fp32_peak_tflops,
    fused_head_propagate_loss,
    fused_propagate_loss,
    fused_propagate_loss_su4,
    philox_errors_su4,
    su4_unitary_generator,
    get_ore_error_distribution,
    get_ore_ple_error_distribution,
    infidelity_loss,
    negative_log_loss,
    philox_errors,
    propagate_fidelity,
    sharp_loss,
    target_coeffs,
    tuning_flags,
)

__version__ = "0.1.0"
from .graphs import GraphedFusedStep  # noqa: F401,E402
from .peer import PeerExchange  # noqa: F401,E402
from .pipeline import PipelinedStep  # noqa: F401,E402
