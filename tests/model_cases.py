"""Shared description of the non-M/M/1 models for the CPU and GPU parity tests."""
from libcxxdes_b200 import abi

MODELS = {
    "aloha": (abi.MODEL_ALOHA, abi.AlohaParams, "aloha_reference.json"),
    "mmc": (abi.MODEL_MMC, abi.MMCParams, "mmc_reference.json"),
    "archsim": (abi.MODEL_ARCHSIM, abi.ArchSimParams, "archsim_reference.json"),
}
N_GOLDEN_CASES = 7


def clock_of(golden):
    return abi.Clock.of(tuple(golden["clock"]["unit"]), tuple(golden["clock"]["precision"]))


def clock_pairs(golden):
    return tuple(golden["clock"]["unit"]), tuple(golden["clock"]["precision"])
