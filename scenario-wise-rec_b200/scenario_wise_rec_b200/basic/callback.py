"""EarlyStopper with the reference's interface (reference: scenario_wise_rec/basic/callback.py:4-33):
keeps a deep copy of the best ``state_dict`` and tells the trainer when ``patience`` validation
rounds passed without a better AUC."""
import copy


class EarlyStopper(object):
    def __init__(self, patience):
        self.patience = patience
        self.trial_counter = 0
        self.best_auc = 0
        self.best_weights = None

    def stop_training(self, val_auc, weights):
        if val_auc > self.best_auc:
            self.best_auc, self.trial_counter = val_auc, 0
            self.best_weights = copy.deepcopy(weights)
            return False
        if self.trial_counter + 1 < self.patience:
            self.trial_counter += 1
            return False
        return True
