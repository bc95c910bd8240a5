"""experiment.py --config --split --l_mix: flag handling and folder mangling (reference experiment.py:31-72,100-111)
on the CPU; a one-epoch end-to-end run (train -> validate -> save -> test) on the GPU."""
import json
import os

import pytest

from multimodal_segmentation_b200.experiment import Experiment


def test_get_config_mangles_folder_like_the_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    a = Experiment.read_console_parameters(["--config", "dafnet_config_chaos", "--split", "2", "--l_mix", "0.5"])
    c = Experiment().get_config(int(a.split), a)
    assert c.folder == "dafnet_chaos_l05_['t1', 't2']_split2"       # '_l%s' % l_mix, modality list, split, '.' stripped
    assert c.l_mix == 0.5 and c.n_pairs == 1 and c.split == 2
    assert c.model == "dafnet.DAFNet" and c.executor == "dafnet_executor.DAFNetExecutor"
    saved = json.load(open(os.path.join(c.folder, "experiment_configuration.json")))
    assert saved["l_mix"] == 0.5 and saved["anatomy_encoder"]["out_channels"] == 8
    a = Experiment.read_console_parameters(["--config", "dafnet_spade_config_chaos", "--split", "0", "--l_mix", "1",
                                            "--automatedpairing", "1", "--randomise", "1"])
    c = Experiment().get_config(0, a)
    assert c.folder.startswith("dafnet_spade_chaos_randomise_automatedpairing_l1_") and c.n_pairs == 3
    assert c.decoder_type == "spade"


def test_l_mix_is_effectively_mandatory(tmp_path, monkeypatch):
    """experiment.py:56-57: hasattr(args, 'l_mix') is always true, so omitting the flag raises float(None)"""
    monkeypatch.chdir(tmp_path)
    a = Experiment.read_console_parameters(["--config", "mmsdnet_config_chaos", "--split", "0"])
    with pytest.raises(TypeError):
        Experiment().get_config(0, a)


def test_required_flags():
    with pytest.raises(SystemExit):
        Experiment.read_console_parameters(["--split", "0"])
    with pytest.raises(SystemExit):
        Experiment.read_console_parameters(["--config", "dafnet_config_chaos"])


@pytest.mark.gpu
def test_experiment_end_to_end_one_epoch(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("DAFK_TRAIN_PAIRS", "8")
    Experiment().run(["--config", "dafnet_config_chaos", "--split", "0", "--l_mix", "1", "--input_size", "64",
                      "--epochs", "1", "--batch_size", "4"])
    folder = "dafnet_chaos_l1_['t1', 't2']_split0"
    rows = open(os.path.join(folder, "training.csv")).read().strip().split("\n")
    assert len(rows) == 2 and rows[0].startswith("epoch,adv_M")
    assert os.path.exists(os.path.join(folder, "test_results_chaos_t2_max", "results.csv"))
