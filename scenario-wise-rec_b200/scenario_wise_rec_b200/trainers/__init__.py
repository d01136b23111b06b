from .ctr_trainer import CTRTrainer

__all__ = ["CTRTrainer"]
