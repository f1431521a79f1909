//go:build cgo

// Package dfr2d is the cgo binding of the B200 device library (include/dfr2d.h) for gocfd's 2D Euler
// DFR time step.  It flattens *Euler2D.Euler into a dfr2d_problem once and then replaces
// `c.RK.Step(c)` (model_problems/Euler2D/euler.go:177).
//
// NOT COMPILED IN THE BUILD CONTAINER (no Go toolchain there).  Written against cmd/cgo rules:
// every pointer handed to C points to Go memory without Go pointers inside and is only used for
// the duration of the call (the library copies during dfr2d_create / dfr2d_set_state).
package dfr2d

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../gocfd_b200/csrc -ldfr2d -Wl,-rpath,${SRCDIR}/../../gocfd_b200/csrc
#include <stdlib.h>
#include "dfr2d.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sort"
	"unsafe"

	"github.com/notargets/gocfd/model_problems/Euler2D"
	"github.com/notargets/gocfd/types"
)

// Solver owns one device partition (n_parts = 1: the whole mesh on one GPU).
type Solver struct {
	h      *C.dfr2d_handle
	np, k  int
	global []float64 // staging for [4][NpInt x K]
}

func d(p []float64) *C.double { return (*C.double)(unsafe.Pointer(&p[0])) }
func i32(p []int32) *C.int32_t { return (*C.int32_t)(unsafe.Pointer(&p[0])) }

func fs(f *Euler2D.FreeStream) (o C.dfr2d_freestream) {
	if f == nil {
		return
	}
	o.Gamma = C.double(f.Gamma)
	for n := 0; n < 4; n++ {
		o.Qinf[n] = C.double(f.Qinf[n])
	}
	o.Pinf, o.QQinf, o.Cinf = C.double(f.Pinf), C.double(f.QQinf), C.double(f.Cinf)
	o.Alpha, o.Minf = C.double(f.Alpha), C.double(f.Minf)
	return
}

// New mirrors the tail of NewEuler + NewRungeKuttaSSP (euler.go:98-117, :343-406).
// kappa is ip.Kappa as parsed; device is the CUDA ordinal.
func New(c *Euler2D.Euler, kappa float64, device int) *Solver {
	return newPartition(c, kappa, 1, 0, device)
}

// newPartition builds partition `part` of `nParts` (PartitionMap.Split1D element ranges) on CUDA device `device`.
func newPartition(c *Euler2D.Euler, kappa float64, nParts, part, device int) *Solver {
	var (
		dfr    = c.DFR
		rt     = dfr.FluxElement
		el     = dfr.SolutionElement
		K      = dfr.K
		NpEdge = rt.NpEdge
	)
	// canonical edge order: (owner element, owner's local edge number) -- Go map order is random
	keys := make([]types.EdgeKey, 0, len(dfr.Tris.Edges))
	for en := range dfr.Tris.Edges {
		keys = append(keys, en)
	}
	sort.Slice(keys, func(a, b int) bool {
		ea, eb := dfr.Tris.Edges[keys[a]], dfr.Tris.Edges[keys[b]]
		ha := int(ea.ConnectedTris[0])*3 + int(ea.ConnectedTriEdgeNumber[0])
		hb := int(eb.ConnectedTris[0])*3 + int(eb.ConnectedTriEdgeNumber[0])
		return ha < hb
	})
	NE := len(keys)
	index := make(map[types.EdgeKey]int32, NE)
	kL, kR := make([]int32, NE), make([]int32, NE)
	numL, numR := make([]int32, NE), make([]int32, NE)
	nconn, bc := make([]int32, NE), make([]int32, NE)
	elen := make([]float64, NE)
	var bpEdge []int32
	var bpX, bpY []float64
	for i, en := range keys {
		e := dfr.Tris.Edges[en]
		index[en] = int32(i)
		kL[i], numL[i] = int32(e.ConnectedTris[0]), int32(e.ConnectedTriEdgeNumber[0])
		nconn[i], bc[i], elen[i] = int32(e.NumConnectedTris), int32(e.BCType), e.GetEdgeLength()
		kR[i] = -1
		if e.NumConnectedTris == 2 {
			kR[i], numR[i] = int32(e.ConnectedTris[1]), int32(e.ConnectedTriEdgeNumber[1])
		} else { // FluxX/FluxY of the edge points (bcs.go:33-37)
			bpEdge = append(bpEdge, int32(i))
			for p := 0; p < NpEdge; p++ {
				row := 2*rt.NpInt + int(numL[i])*NpEdge + p
				bpX = append(bpX, dfr.FluxX.DataP[int(kL[i])+row*K])
				bpY = append(bpY, dfr.FluxY.DataP[int(kL[i])+row*K])
			}
		}
	}
	etoEdge := make([]int32, 3*K)
	etov := make([]int32, 3*K)
	for k := 0; k < K; k++ {
		for e := 0; e < 3; e++ {
			etoEdge[3*k+e] = index[dfr.EdgeNumber[k+K*e]]
			etov[3*k+e] = int32(dfr.Tris.EToV.DataP[3*k+e])
		}
	}
	if len(bpEdge) == 0 { // keep &x[0] valid
		bpEdge, bpX, bpY = []int32{0}, []float64{0}, []float64{0}
	}
	sf := c.RK.ShockSensor[0]
	var p C.dfr2d_problem
	p.N = C.int32_t(dfr.N)
	p.flux_type, p.init_case = C.int32_t(c.FluxCalcAlgo), C.int32_t(c.Case)
	if c.LocalTimeStepping {
		p.local_time_stepping = 1
	}
	p.max_iterations = C.int32_t(c.MaxIterations)
	var bary []float64
	if c.Dissipation != nil {
		p.dissipation = 1
		bary = c.Dissipation.BaryCentricCoords.DataP
	} else {
		bary = make([]float64, rt.Np*3)
	}
	p.K, p.NV, p.NE = C.int64_t(K), C.int64_t(dfr.VX.Len()), C.int64_t(NE)
	p.NBP = C.int64_t(len(bpX) / NpEdge)
	p.CFL, p.FinalTime, p.Kappa = C.double(c.CFL), C.double(c.FinalTime), C.double(kappa)
	p.FSFar, p.FSIn, p.FSOut = fs(c.FSFar), fs(c.FSIn), fs(c.FSOut)
	p.vortex = C.dfr2d_vortex{Beta: 5, X0: 5, Y0: 0, Gamma: 1.4, Ufs: 1} // InitializeIVortex, initialization.go:61-66
	p.FluxEdgeInterp, p.DivInt, p.Div = d(dfr.FluxEdgeInterp.DataP), d(rt.DivInt.DataP), d(rt.Div.DataP)
	p.V, p.Vinv = d(el.JB2D.V.DataP), d(el.JB2D.Vinv.DataP)
	p.MassMatrix, p.D, p.P, p.ModeFilter = d(sf.MassMatrix.DataP), d(sf.D.DataP), d(sf.P.DataP), d(sf.ModeFilter)
	p.Bary = d(bary)
	p.Jdet, p.Jinv = d(dfr.Jdet.DataP), d(dfr.Jinv.DataP)
	p.FaceNormX, p.FaceNormY = d(dfr.FaceNorm[0].DataP), d(dfr.FaceNorm[1].DataP)
	p.IInII, p.EdgeLenMax = d(dfr.IInII.DataP), d(dfr.EdgeLenMax.DataP)
	p.EToV, p.EtoEdge = i32(etov), i32(etoEdge)
	p.edge_kL, p.edge_kR, p.edge_numL, p.edge_numR = i32(kL), i32(kR), i32(numL), i32(numR)
	p.edge_nconn, p.edge_bc, p.edge_len = i32(nconn), i32(bc), d(elen)
	p.bp_edge, p.bp_x, p.bp_y = i32(bpEdge), d(bpX), d(bpY)

	s := &Solver{np: el.Np, k: K, global: make([]float64, 4*el.Np*K)}
	// C.dfr2d_problem holds Go pointers: pin them for the call (Go >= 1.21)
	var pin runtime.Pinner
	defer pin.Unpin()
	for _, ptr := range []unsafe.Pointer{unsafe.Pointer(p.FluxEdgeInterp), unsafe.Pointer(p.DivInt), unsafe.Pointer(p.Div),
		unsafe.Pointer(p.V), unsafe.Pointer(p.Vinv), unsafe.Pointer(p.MassMatrix), unsafe.Pointer(p.D), unsafe.Pointer(p.P),
		unsafe.Pointer(p.ModeFilter), unsafe.Pointer(p.Bary), unsafe.Pointer(p.Jdet), unsafe.Pointer(p.Jinv),
		unsafe.Pointer(p.FaceNormX), unsafe.Pointer(p.FaceNormY), unsafe.Pointer(p.IInII), unsafe.Pointer(p.EdgeLenMax),
		unsafe.Pointer(p.EToV), unsafe.Pointer(p.EtoEdge), unsafe.Pointer(p.edge_kL), unsafe.Pointer(p.edge_kR),
		unsafe.Pointer(p.edge_numL), unsafe.Pointer(p.edge_numR), unsafe.Pointer(p.edge_nconn), unsafe.Pointer(p.edge_bc),
		unsafe.Pointer(p.edge_len), unsafe.Pointer(p.bp_edge), unsafe.Pointer(p.bp_x), unsafe.Pointer(p.bp_y)} {
		pin.Pin(ptr)
	}
	if rc := C.dfr2d_create(&p, C.int(nParts), C.int(part), C.int(device), &s.h); rc != 0 {
		panic(fmt.Errorf("dfr2d_create: %s", C.GoString(C.dfr2d_last_error(nil))))
	}
	s.SetState(c)
	return s
}

func (s *Solver) check(rc C.int, what string) {
	if rc != 0 {
		panic(fmt.Errorf("%s: %s", what, C.GoString(C.dfr2d_last_error(s.h)))) // "NAN found" etc.
	}
}

// SetState uploads c.Q (RecombineShardsKBy4 layout, parallelism.go:86-98).
func (s *Solver) SetState(c *Euler2D.Euler) {
	c.RecombineShardsKBy4(c.Q, &c.Q4)
	for n := 0; n < 4; n++ {
		copy(s.global[n*s.np*s.k:(n+1)*s.np*s.k], c.Q4[n].DataP)
	}
	s.check(C.dfr2d_set_state(s.h, d(s.global)), "dfr2d_set_state")
}

// GetState downloads into c.Q4 (the un-sharded solution the plot/output code reads).
func (s *Solver) GetState(c *Euler2D.Euler) {
	s.check(C.dfr2d_get_state(s.h, d(s.global)), "dfr2d_get_state")
	for n := 0; n < 4; n++ {
		copy(c.Q4[n].DataP, s.global[n*s.np*s.k:(n+1)*s.np*s.k])
	}
}

// Step runs nsteps x {RK.Step; Time += GlobalDT; StepCount++} and returns what Solve needs.
func (s *Solver) Step(nsteps int) (time, dt float64, steps int, finished bool) {
	var info C.dfr2d_step_info
	s.check(C.dfr2d_step(s.h, C.int(nsteps), &info), "dfr2d_step")
	return float64(info.time), float64(info.dt), int(info.steps), info.finished != 0
}

// Residual is the per-variable signed max PrintUpdate prints (euler.go:821-835).
func (s *Solver) Residual() (r [4]float64) {
	s.check(C.dfr2d_residual(s.h, (*C.double)(unsafe.Pointer(&r[0]))), "dfr2d_residual")
	return
}

// InitState evaluates Euler.InitializeSolution (euler.go:728-794) on the device instead of uploading c.Q.
func (s *Solver) InitState(c *Euler2D.Euler) {
	dfr := c.DFR
	etov := make([]int32, 3*s.k)
	for k := 0; k < s.k; k++ {
		for v := 0; v < 3; v++ {
			etov[3*k+v] = int32(dfr.Tris.EToV.At(k, v))
		}
	}
	el := dfr.SolutionElement
	s.check(C.dfr2d_init_state(s.h, C.int(c.Case), C.int64_t(dfr.VX.Len()), d(dfr.VX.DataP), d(dfr.VY.DataP), i32(etov),
		d(el.R.DataP), d(el.S.DataP)), "dfr2d_init_state")
}

// PlotField is Euler.GetPlotField (plot.go:14-86) for the GetFlowFunction family, evaluated on the device:
// flow function -> GraphInterp -> AverageGraphFieldVertices -> transpose -> float32, i.e. exactly the slice
// AVSFieldWriter.saveField stores (DG2D/graphics_support.go:80-93).  The caller hands it to the writer instead of
// RecombineShardsKBy4 + GetPlotField (euler.go:192-207).
func (s *Solver) PlotField(c *Euler2D.Euler, ff Euler2D.FlowFunction) []float32 {
	gi := c.DFR.GraphInterp
	nr, _ := gi.Dims()
	out := make([]float32, s.k*nr)
	s.check(C.dfr2d_plot_field(s.h, C.int(ff), d(gi.DataP), C.int(nr), (*C.float)(unsafe.Pointer(&out[0]))), "dfr2d_plot_field")
	return out
}

// CaptureEdgeValues keeps the EdgeQValues store (the owner side's Q_Face of every step's last stage, edges.go:344-350)
// on the device from the next step on; GradientField needs it.  Turn it on before the step whose fields are plotted.
func (s *Solver) CaptureEdgeValues(c *Euler2D.Euler, on bool) {
	if !on {
		s.check(C.dfr2d_capture_edge_values(s.h, 0, nil, nil), "dfr2d_capture_edge_values")
		return
	}
	fn := c.DFR.FaceNorm
	s.check(C.dfr2d_capture_edge_values(s.h, 1, d(fn[0].DataP), d(fn[1].DataP)), "dfr2d_capture_edge_values")
}

// GradientField replaces GetPlotField for XGradientDensity..YGradientEnergy (plot.go:54-77): [NpFlux x K] doubles, not
// interpolated, as GetSolutionGradientUsingRTElement(-1, n, c.Q, ...) leaves GradX / GradY.
func (s *Solver) GradientField(c *Euler2D.Euler, ff Euler2D.FlowFunction) []float64 {
	out := make([]float64, c.DFR.FluxElement.Np*s.k)
	s.check(C.dfr2d_gradient_field(s.h, C.int(ff), d(out)), "dfr2d_gradient_field")
	return out
}

// HilbertOrder returns order[new] = old element along a Hilbert curve through the rank-normalised element centroids: apply
// it to EToV right after the mesh is read (before NewDFR2D) so that the contiguous PartitionMap ranges are spatially compact.
func HilbertOrder(EToV []int32, VX, VY []float64) []int32 {
	order := make([]int32, len(EToV)/3)
	if rc := C.dfr2d_hilbert_order(C.int64_t(len(order)), C.int64_t(len(VX)), i32(EToV), d(VX), d(VY), i32(order)); rc != 0 {
		panic(fmt.Errorf("dfr2d_hilbert_order: %s", C.GoString(C.dfr2d_last_error(nil))))
	}
	return order
}

func (s *Solver) Close() { C.dfr2d_destroy(s.h); s.h = nil }

// MultiSolver drives one handle per GPU from a single Go process -- the shape of the reference's controller goroutine
// (euler.go:408-412): dfr2d_multi_step runs the per-stage protocol of include/dfr2d.h over all partitions.  The partitions
// exchange Q_Face halo rows, shared-vertex maxima, DissX/DissY edge rows and the wave-speed maxima among themselves:
// pack kernels store straight into the partner GPU's mailbox over NVLink and raise an arrival flag, unpack kernels wait
// on the flag (gocfd_b200/csrc/dfr2d_peer.cuh) -- no host synchronisation, no collective library.  Measured on 2 B200:
// 1.39e11 DOF-stage-updates/s at config C5, bitwise equal to one partition (profiles/r02b_*).  (One process per GPU
// uses dfr2d_peer_export / dfr2d_peer_connect and then plain dfr2d_step on every partition.)
type MultiSolver struct {
	Parts []*Solver
}

// NewMulti creates one partition per device (devices[g] may repeat: several partitions per GPU).
func NewMulti(c *Euler2D.Euler, kappa float64, devices []int) *MultiSolver {
	m := &MultiSolver{}
	for g, dev := range devices {
		m.Parts = append(m.Parts, newPartition(c, kappa, len(devices), g, dev))
	}
	return m
}

// Step runs nsteps x {RK.Step; Time += GlobalDT; StepCount++} over all partitions.
func (m *MultiSolver) Step(nsteps int) (time, dt float64, steps int, finished bool) {
	hs := make([]*C.dfr2d_handle, len(m.Parts))
	for g, s := range m.Parts {
		hs[g] = s.h
	}
	var info C.dfr2d_step_info
	if rc := C.dfr2d_multi_step(&hs[0], C.int(len(hs)), C.int(nsteps), &info); rc != 0 {
		for _, s := range m.Parts {
			s.check(rc, "dfr2d_multi_step")
		}
	}
	return float64(info.time), float64(info.dt), int(info.steps), info.finished != 0
}

func (m *MultiSolver) handles() []*C.dfr2d_handle {
	hs := make([]*C.dfr2d_handle, len(m.Parts))
	for g, s := range m.Parts {
		hs[g] = s.h
	}
	return hs
}

// SetState scatters c.Q4 to all partitions; the copies to the different GPUs cross PCIe concurrently.
func (m *MultiSolver) SetState(c *Euler2D.Euler) {
	g := m.Parts[0]
	for n := 0; n < 4; n++ {
		copy(g.global[n*g.np*g.k:(n+1)*g.np*g.k], c.Q4[n].DataP)
	}
	hs := m.handles()
	g.check(C.dfr2d_multi_set_state(&hs[0], C.int(len(hs)), d(g.global)), "dfr2d_multi_set_state")
}

// GetState gathers every partition's columns into c.Q4 (each handle writes only its own element range).
func (m *MultiSolver) GetState(c *Euler2D.Euler) {
	hs := m.handles()
	m.Parts[0].check(C.dfr2d_multi_get_state(&hs[0], C.int(len(hs)), d(m.Parts[0].global)), "dfr2d_multi_get_state")
	g := m.Parts[0]
	for n := 0; n < 4; n++ {
		copy(c.Q4[n].DataP, g.global[n*g.np*g.k:(n+1)*g.np*g.k])
	}
}

// Residual is the per-variable signed max over all partitions (PrintUpdate's loop over np, euler.go:823-829).
func (m *MultiSolver) Residual() (r [4]float64) {
	hs := m.handles()
	m.Parts[0].check(C.dfr2d_multi_residual(&hs[0], C.int(len(hs)), (*C.double)(unsafe.Pointer(&r[0]))), "dfr2d_multi_residual")
	return
}

func (m *MultiSolver) Close() {
	for _, s := range m.Parts {
		s.Close()
	}
}
