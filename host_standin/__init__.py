"""Stand-ins for parts of the Go host that stay in Go and never cross the C ABI (OUT OF SCOPE):
OutputFinal's post-processing (euler.go:221-318) and the Sod shock-tube sampler / exact solution
(model_problems/Euler1D/sod_shock_tube).  Test infrastructure for the validation cases (C3 profile checks) -- not part of
the product package gocfd_b200/, which is the device library, its binding and the mirror of the interface it consumes."""
