"""scenario_wise_rec_b200 -- B200-native drop-in for the Scenario-Wise-Rec hot path.

Mirrors the reference package layout for the path in scope:

    scenario_wise_rec.basic.features      -> scenario_wise_rec_b200.basic.features
    scenario_wise_rec.basic.layers        -> scenario_wise_rec_b200.basic.layers
    scenario_wise_rec.models.multi_domain -> scenario_wise_rec_b200.models.multi_domain
    scenario_wise_rec.trainers.CTRTrainer -> scenario_wise_rec_b200.trainers.CTRTrainer

Same constructor signatures, ``forward(x_dict) -> Tensor[B]``, parameter names and
``state_dict`` keys; the compute is hand-written sm_100a CUDA behind the C ABI in
``include/swr_b200.h`` (``libswr_b200.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
