"""Test-set evaluation (reference: model_tester.py:12-85): per-volume Dice of the binarised prediction for the
'simple', 'def' and 'max' prediction types, written to <folder>/test_results_<dataset>_<modality>_<type>/results.csv.
The PNG dumps (scipy.misc.imsave, model_tester.py:87-102) are dropped; pair randomisation belongs to the
automated-pairing rows and is skipped."""
import logging
import os

import numpy as np

from . import costs
from .loaders import loader_factory

log = logging.getLogger("model_tester")


class ModelTester(object):
    def __init__(self, model, conf):
        self.model = model
        self.conf = conf

    def run(self):
        for modi, mod in enumerate(self.model.modalities):
            log.info("Evaluating model on test data for %s" % mod)
            self.test_modality(mod, modi)

    def make_test_folder(self, modality, suffix=""):
        folder = os.path.join(self.conf.folder, "test_results_%s_%s_%s" % (self.conf.test_dataset, modality, suffix))
        if not os.path.exists(folder):
            os.makedirs(folder)
        return folder

    def test_modality(self, modality, modality_index):
        test_loader = loader_factory.init_loader(self.conf.test_dataset)
        test_loader.modalities = self.conf.modality
        if hasattr(test_loader, "input_shape"):
            test_loader.input_shape = tuple(self.conf.input_shape)
        test_data = test_loader.load_all_modalities_concatenated(self.conf.split, "test", self.conf.image_downsample)
        results = {}
        for type in ["simple", "def", "max"]:
            folder = self.make_test_folder(modality, suffix=type)
            results[type] = self.test_modality_type(folder, modality_index, type, test_loader, test_data)
        return results

    def test_modality_type(self, folder, modality_index, type, test_loader, test_data):
        assert type in ["simple", "def", "max", "maxnostn"]
        nm = test_loader.num_masks
        im_dice = {}
        with open(os.path.join(folder, "results.csv"), "w") as f:
            f.writelines("Vol, Dice, " + ", ".join(["Dice%d" % mi for mi in range(nm)]) + "\n")
            for vol_i, sl in _volumes(test_data):
                vol_image_mod1 = test_data.get_images_modi(0)[sl]
                vol_image_mod2 = test_data.get_images_modi(1)[sl]
                assert vol_image_mod1.shape[0] > 0
                vol_mask = test_data.get_masks_modi(modality_index)[sl][..., :nm]
                prd_mask = self.model.predict_mask(modality_index, type, [vol_image_mod1, vol_image_mod2])
                im_dice[vol_i] = costs.dice(vol_mask, prd_mask, binarise=True)
                sep_dice = [costs.dice(vol_mask[..., mi:mi + 1], prd_mask[..., mi:mi + 1], binarise=True) for mi in range(nm)]
                s = "%s, %.3f, " + ", ".join(["%.3f"] * nm) + "\n"
                f.writelines(s % ((str(vol_i), im_dice[vol_i]) + tuple(sep_dice)))
        score = float(np.mean(list(im_dice.values())))
        print("%s - Dice score: %.3f" % (type, score))
        return score


def _volumes(data, slices_per_volume=8):
    """synthetic data has no volume ids: consecutive groups of slices stand in for volumes"""
    n = data.size()
    for v, a in enumerate(range(0, n, slices_per_volume)):
        yield v, slice(a, min(n, a + slices_per_volume))
