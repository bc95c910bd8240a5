"""Entry point (reference: experiment.py:16-129):

    python experiment.py --config dafnet_config_chaos --split 0 --l_mix 1 [--test 1] [--test_dataset chaos]
                         [--automatedpairing 1] [--randomise 1]

Same flags, same folder-name mangling, same `config.model` / `config.executor` dotted-name dispatch.  GitPython,
comet_ml and matplotlib are dropped (they never touch the step).  Extras for synthetic runs, all optional:
`--input_size N` (square slices of N pixels, must be divisible by 32), `--epochs`, `--batch_size`.
"""
import argparse
import importlib
import json
import logging
import os

import numpy

from .keras_like import EasyDict


class Experiment(object):
    def __init__(self):
        self.log = None

    def init_logging(self, config):
        if not os.path.exists(config.folder):
            os.makedirs(config.folder)
        logging.basicConfig(filename=config.folder + "/logfile.log", level=logging.DEBUG, format="%(asctime)s %(message)s")
        logging.getLogger().addHandler(logging.StreamHandler())
        self.log = logging.getLogger()
        self.log.debug(config.items())
        self.log.info("---- Setting up experiment at " + config.folder + "----")

    def get_config(self, split, args):
        """experiment.py:31-72"""
        mod = importlib.import_module("multimodal_segmentation_b200.configuration." + args.config)
        size = getattr(args, "input_size", None)
        config_dict = mod.get((int(size), int(size), 1)) if size else mod.get()
        config = EasyDict(config_dict)
        config.split = split
        if (hasattr(config, "randomise") and config.randomise) or (hasattr(args, "randomise") and args.randomise):
            config.randomise = True
            config.folder += "_randomise"
        config.n_pairs = 1
        if (hasattr(config, "automatedpairing") and config.automatedpairing) or \
                (hasattr(args, "automatedpairing") and args.automatedpairing):
            config.automatedpairing = True
            config.folder += "_automatedpairing"
            config.n_pairs = 3
        l_mix = config.l_mix
        if hasattr(args, "l_mix"):
            # as in the reference, argparse always defines the attribute: omitting --l_mix raises here
            config.l_mix = float(args.l_mix)
            l_mix = args.l_mix
        config.folder += "_l%s" % l_mix
        config.folder += "_" + str(config.modality)
        config.folder += "_split%s" % split
        config.folder = config.folder.replace(".", "")
        if args.test_dataset:
            print("Overriding default test dataset")
            config.test_dataset = args.test_dataset
        if getattr(args, "epochs", None):
            config.epochs = int(args.epochs)
        if getattr(args, "batch_size", None):
            config.batch_size = int(args.batch_size)
        config.githash = "n/a"
        self.save_config(config)
        return config

    def save_config(self, config):
        if not os.path.exists(config.folder):
            os.makedirs(config.folder)
        with open(config.folder + "/experiment_configuration.json", "w") as outfile:
            json.dump(_plain(config), outfile)

    def run(self, argv=None):
        args = Experiment.read_console_parameters(argv)
        configuration = self.get_config(int(args.split), args)
        self.init_logging(configuration)
        self.run_experiment(configuration, args.test)

    def run_experiment(self, configuration, test):
        executor = self.get_executor(configuration, test)
        if test:
            executor.test()
        else:
            executor.train()
            with open(configuration.folder + "/experiment_configuration.json", "w") as outfile:
                json.dump(_plain(configuration), outfile)
            executor.test()

    @staticmethod
    def read_console_parameters(argv=None):
        parser = argparse.ArgumentParser(description="")
        parser.add_argument("--config", default="", help="The experiment configuration file", required=True)
        parser.add_argument("--test", help="Evaluate the model on test data", type=bool)
        parser.add_argument("--test_dataset", help="Override default test dataset", choices=["chaos"])
        parser.add_argument("--split", help="Data split to run.", required=True)
        parser.add_argument("--l_mix", help="Percentage of labelled data")
        parser.add_argument("--automatedpairing", help="Use weighted cost for training", type=bool)
        parser.add_argument("--randomise", help="Randomise multimodal pairs", type=bool)
        parser.add_argument("--input_size", help="(synthetic data) slice size in pixels", type=int)
        parser.add_argument("--epochs", help="override the number of epochs", type=int)
        parser.add_argument("--batch_size", help="override the batch size", type=int)
        return parser.parse_args(argv)

    def get_executor(self, config, test):
        """experiment.py:113-124"""
        module_name, model_name = config.model.split(".")[0], config.model.split(".")[1]
        model = getattr(importlib.import_module("multimodal_segmentation_b200.models." + module_name), model_name)(config)
        model.build()
        module_name, model_name = config.executor.split(".")[0], config.executor.split(".")[1]
        executor = getattr(importlib.import_module("multimodal_segmentation_b200.model_executors." + module_name),
                           model_name)(config, model)
        return executor


def _plain(o):
    if isinstance(o, dict):
        return {k: _plain(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_plain(v) for v in o]
    if isinstance(o, numpy.integer):
        return int(o)
    if isinstance(o, numpy.floating):
        return float(o)
    return o


if __name__ == "__main__":
    Experiment().run()
