"""Command-line configuration with the reference's option names and defaults (reference: config.py:13-83), so that
`python train.py --experiment TEDM ...` lines written for the reference keep working."""
from __future__ import annotations

import argparse
import os
from datetime import datetime

this_dir = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
default_logdir = os.path.join(this_dir, "logs", datetime.now().strftime("%Y%m%d_%H%M%S"))

EXPERIMENTS = ["img_only", "PDDM", "baseline", "LEDM", "LEDMe", "TEDM"]
# the contrastive baselines (global_cl, local_cl, global_finetune, glob_loc_finetune) are outside the path

parser = argparse.ArgumentParser()
parser.add_argument("--debug", action="store_true")
parser.add_argument("--mixed_precision", type=bool, default=False,
                    help="accepted for compatibility; the kernels always compute in bf16 with fp32 accumulation")
parser.add_argument("--resume_path", type=str, default=None, help="Path to checkpoint to resume from")
# Experiment parameters
parser.add_argument("--experiment", type=str, default="img_only", choices=EXPERIMENTS)
parser.add_argument("--dataset", type=str, default="JSRT", choices=["JSRT", "CXR14", "synthetic"], help="Dataset to use")
# Data parameters
parser.add_argument("--img_size", type=int, default=128, help="Height / width of the input image to the network")
parser.add_argument("--data_dir", type=str, help="Path to the dataset")
parser.add_argument("--csv_dir", type=str, default=None,
                    help="Directory holding the split CSVs (train_split.csv, JSRT_{train,val,test}_split.csv); default <repo>/data "
                         "as in the reference, which ships them in its own data/ directory")
parser.add_argument("--num_workers", type=int, default=4, help="Number of subprocesses to use for data loading")
# Model parameters
parser.add_argument("--dim", type=int, default=64, help="Width of the U-Net")
parser.add_argument("--dim_mults", nargs="+", type=int, default=(1, 2, 4, 8), help="Dimension multipliers for U-Net levels")
# Diffusion parameters
parser.add_argument("--timesteps", type=int, default=1000, help="Number of diffusion timesteps")
parser.add_argument("--beta_schedule", type=str, default="cosine", choices=["linear", "cosine"])
parser.add_argument("--objective", type=str, default="pred_noise", choices=["pred_noise", "pred_x_0"])
# Training parameters
parser.add_argument("--batch_size", type=int, default=16, help="Input batch size (per GPU)")
parser.add_argument("--lr", type=float, default=1e-4, help="Learning rate")
parser.add_argument("--weight_decay", type=float, default=0, help="Weight decay")
parser.add_argument("--max_steps", type=int, default=500000, help="Number of training steps to perform")
parser.add_argument("--p2_loss_weight_gamma", type=float, default=0.)
parser.add_argument("--p2_loss_weight_k", type=float, default=1.)
parser.add_argument("--device", type=str, default="cuda", help="cuda only: there is no CPU fallback")
parser.add_argument("--seed", type=int, default=0, help="Random seed")
# Logging parameters
parser.add_argument("--log_freq", type=int, default=100, help="Frequency of logging")
parser.add_argument("--val_freq", type=int, default=100, help="Frequency of validation")
parser.add_argument("--val_steps", type=int, default=250, help="Number of timestep to use for validation")
parser.add_argument("--log_dir", type=str, default=default_logdir, help="Logging directory")
parser.add_argument("--n_sampled_imgs", type=int, default=8, help="Number of images to sample during logging")
parser.add_argument("--max_val_steps", type=int, default=-1, help="Number of validation steps to perform")
# datasetGAN like segmentation model parameters
parser.add_argument("--saved_diffusion_model", type=str, default="logs/20230127_164150/best_model.pt",
                    help="Path to checkpoint of trained diffusion model")
parser.add_argument("--t_steps_to_save", type=int, nargs="*", choices=range(1000), default=[50, 200, 400, 600, 800],
                    help="Diffusion steps to be used as features")
parser.add_argument("--n_labelled_images", type=int, default=None, choices=[197, 98, 49, 24, 12, 6, 3, 1],
                    help="Number of labelled images to use for semi-supervised training")
parser.add_argument("--shared_weights_over_timesteps", default=False, action="store_true")
parser.add_argument("--early_stop", default=False, action="store_true")
# additions of this implementation
parser.add_argument("--no_cuda_graph", dest="cuda_graph", action="store_false", help="run the DDPM step eagerly")
parser.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"],
                    help="bf16: bf16 storage / fp32 accumulation (training and inference); fp32: the reference's default "
                         "arithmetic to 1e-4 through split-bf16 tensor-core convolutions (inference only)")
parser.add_argument("--sync_bn", default=False, action="store_true",
                    help="data-parallel head training: share the head's BatchNorm batch statistics over all ranks, so N GPUs "
                         "with B/N images each equal the single-device reference at batch B (default: per-replica statistics)")
