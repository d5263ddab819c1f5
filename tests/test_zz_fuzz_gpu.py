"""The randomised collective sequences of test_fuzz_cpu.py on the CUDA kernels (loop-back ranks on one GPU).  Kept in a
last-sorted file: its first hardware run happens at round end."""
import pytest

from test_fuzz_cpu import run_fuzz

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,seed", [(4, 2), (6, 4)])
def test_random_collective_sequences_device(world, seed):
    run_fuzz(world, seed, "cuda")
