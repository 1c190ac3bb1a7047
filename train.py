"""Entry point with the reference's command line (reference: train.py:15-56).

    python train.py --experiment {img_only,baseline,LEDM,LEDMe,TEDM} [--dataset synthetic] ...
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py --experiment img_only ...

One process per GPU; under torchrun every rank trains on its own shard of the data and the gradients are
summed with one NCCL all-reduce per step."""
import argparse
from pathlib import Path

from tedm_b200.config import parser


def main(argv=None):
    config = argparse.ArgumentParser(parents=[parser], add_help=False).parse_args(argv)
    config.normalize = True
    config.log_dir = Path(config.log_dir).parent / config.experiment / str(config.n_labelled_images) / Path(config.log_dir).name
    config.channels = 1
    config.out_channels = 1
    if config.data_dir is not None:
        config.data_dir = Path(config.data_dir)
    if config.experiment in ("img_only", "PDDM"):
        from tedm_b200.trainers.train_CXR14 import main as run
    elif config.experiment == "baseline":
        from tedm_b200.trainers.train_baseline import main as run
    elif config.experiment == "LEDM":
        config.t_steps_to_save = [50, 150, 250]
        from tedm_b200.trainers.train_datasetDM import main as run
    elif config.experiment == "LEDMe":
        config.t_steps_to_save = [1, 10, 25, 50, 200, 400, 600, 800]
        from tedm_b200.trainers.train_datasetDM import main as run
    elif config.experiment == "TEDM":
        config.shared_weights_over_timesteps = True
        config.t_steps_to_save = [1, 10, 25, 50, 200, 400, 600, 800]
        from tedm_b200.trainers.train_datasetDM import main as run
    else:
        raise ValueError(f"Unknown experiment: {config.experiment}")
    run(config)


if __name__ == "__main__":
    main()
