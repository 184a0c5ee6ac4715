"""Host-side mirror of the reference's src/ package (environment, history, ops,
network, agent), driving the sm_100a kernels through the C-ABI."""
