"""Host helpers the reference's eval / serve / train scripts import from `groma.utils` (reference `groma/utils.py`):
`disable_torch_init` (:92-98), `init_distributed_mode` + `setup_for_distributed` (:128-185), the logging helpers the
serve layer imports (`build_logger`, `StreamToLogger`, :16-89, `pretty_print_semaphore` :122-125) and the two message
constants.  Behaviour and names follow the reference so `groma/eval/*.py` run unchanged on top of this package; the
process-group backend stays NCCL (one process per GPU)."""
from __future__ import annotations

import logging
import logging.handlers
import os
import subprocess
import sys

import torch

from groma.constants import LOGDIR

server_error_msg = "**NETWORK ERROR DUE TO HIGH TRAFFIC. PLEASE REGENERATE OR REFRESH THIS PAGE.**"
moderation_msg = "YOUR INPUT VIOLATES OUR CONTENT MODERATION GUIDELINES. PLEASE TRY AGAIN."

handler = None


class StreamToLogger:
    """File-like object forwarding complete lines to a logger (the serve layer swaps sys.stdout / sys.stderr for it)."""

    def __init__(self, logger, log_level=logging.INFO):
        self.terminal = sys.stdout
        self.logger = logger
        self.log_level = log_level
        self.linebuf = ""

    def __getattr__(self, attr):
        return getattr(self.terminal, attr)

    def write(self, buf):
        pending, self.linebuf = self.linebuf + buf, ""
        for line in pending.splitlines(True):
            if line.endswith("\n"):
                self.logger.log(self.log_level, line.rstrip())
            else:
                self.linebuf += line

    def flush(self):
        if self.linebuf:
            self.logger.log(self.log_level, self.linebuf.rstrip())
        self.linebuf = ""


def build_logger(logger_name, logger_filename):
    """Root formatter + stdout/stderr redirection + one daily-rotating file handler under LOGDIR shared by all loggers."""
    global handler
    formatter = logging.Formatter(fmt="%(asctime)s | %(levelname)s | %(name)s | %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    if not logging.getLogger().handlers:
        logging.basicConfig(level=logging.INFO)
    logging.getLogger().handlers[0].setFormatter(formatter)
    for name, level, attr in (("stdout", logging.INFO, "stdout"), ("stderr", logging.ERROR, "stderr")):
        lg = logging.getLogger(name)
        lg.setLevel(level)
        setattr(sys, attr, StreamToLogger(lg, level))
    logger = logging.getLogger(logger_name)
    logger.setLevel(logging.INFO)
    if handler is None:
        os.makedirs(LOGDIR, exist_ok=True)
        handler = logging.handlers.TimedRotatingFileHandler(os.path.join(LOGDIR, logger_filename), when="D", utc=True)
        handler.setFormatter(formatter)
        for item in logging.root.manager.loggerDict.values():
            if isinstance(item, logging.Logger):
                item.addHandler(handler)
    return logger


def disable_torch_init():
    """Skip the default nn.Linear / nn.LayerNorm initialisers (weights are always loaded afterwards)."""
    setattr(torch.nn.Linear, "reset_parameters", lambda self: None)
    setattr(torch.nn.LayerNorm, "reset_parameters", lambda self: None)


def violates_moderation(text):
    """OpenAI moderation probe used by the (stale) gradio server; any transport / schema failure means "not flagged"."""
    import requests
    headers = {"Content-Type": "application/json", "Authorization": "Bearer " + os.environ["OPENAI_API_KEY"]}
    data = ("{" + '"input": ' + f'"{text.replace(chr(10), "")}"' + "}").encode("utf-8")
    try:
        ret = requests.post("https://api.openai.com/v1/moderations", headers=headers, data=data, timeout=5)
        return ret.json()["results"][0]["flagged"]
    except (requests.exceptions.RequestException, KeyError):
        return False


def pretty_print_semaphore(semaphore):
    if semaphore is None:
        return "None"
    return f"Semaphore(value={semaphore._value}, locked={semaphore.locked()})"


def setup_for_distributed(is_master):
    """After this call `print` is silent on non-master ranks unless called with force=True."""
    import builtins
    builtin_print = builtins.print

    def print(*args, **kwargs):  # noqa: A001 - the reference replaces the builtin on purpose
        force = kwargs.pop("force", False)
        if is_master or force:
            builtin_print(*args, **kwargs)

    builtins.print = print


def init_distributed_mode(args):
    """Fill args.{rank, world_size, gpu, dist_url, distributed, dist_backend} from the torchrun or SLURM environment, bind
    the process to its GPU and join the NCCL process group (reference groma/utils.py:143-185; callers:
    eval/eval_rec.py:146, eval/model_vg.py, eval/model_refcocog.py)."""
    env = os.environ
    if "RANK" in env and "WORLD_SIZE" in env:
        args.rank, args.world_size, args.gpu = int(env["RANK"]), int(env["WORLD_SIZE"]), int(env["LOCAL_RANK"])
        args.dist_url = "env://"
        env["LOCAL_SIZE"] = str(torch.cuda.device_count())
        print("Using distributed mode: 1")
    elif "SLURM_PROCID" in env:
        proc_id, ntasks = int(env["SLURM_PROCID"]), int(env["SLURM_NTASKS"])
        num_gpus = torch.cuda.device_count()
        addr = subprocess.getoutput("scontrol show hostname {} | head -n1".format(env["SLURM_NODELIST"]))
        env["MASTER_PORT"] = env.get("MASTER_PORT", "29500")
        env["MASTER_ADDR"] = addr
        env["WORLD_SIZE"], env["RANK"] = str(ntasks), str(proc_id)
        env["LOCAL_RANK"], env["LOCAL_SIZE"] = str(proc_id % num_gpus), str(num_gpus)
        args.dist_url = "env://"
        args.world_size, args.rank, args.gpu = ntasks, proc_id, proc_id % num_gpus
        print("Using distributed mode: slurm")
        print(f"world: {env['WORLD_SIZE']}, rank:{env['RANK']}, local_rank{env['LOCAL_RANK']}, local_size{env['LOCAL_SIZE']}")
    else:
        print("Not using distributed mode")
        args.distributed = False
        return
    args.distributed = True
    torch.cuda.set_device(args.gpu)
    args.dist_backend = "nccl"
    print("| distributed init (rank {}): {}".format(args.rank, args.dist_url), flush=True)
    torch.distributed.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                                         rank=args.rank)
    torch.distributed.barrier()
    setup_for_distributed(args.rank == 0)
