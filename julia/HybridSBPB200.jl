# HybridSBPB200.jl -- thin `ccall` layer over libhsbp.so (include/hsbp.h) plus the drop-in type for the
# reference's `factorization` plugin.
#
# NOT EXECUTED IN THE BUILD CONTAINER: Julia is not installed there (tests/test_binding_signatures_cpu.py checks every ccall's
# symbol, return type and argument types, and the struct layouts, against include/hsbp.h).  This file mirrors, call for call,
# hybridsbp_b200/_lib.py + blocks.py (the ctypes twin that the tests run); keep the two in sync.
#
# Usage inside the reference (square_circle.jl:297-299, seas/BP1/BP1.jl:78):
#
#     include("global_curved.jl"); include("HybridSBPB200.jl"); using .HybridSBPB200
#     ctx = HybridSBPB200.Context(0)
#     # metrics as create_metrics (global_curved.jl:136-209) returns them, LFToB as locoperator takes it
#     blk = HybridSBPB200.Blocks(ctx, SBPp, Nr, Ns)                    # Nr, Ns :: Vector{Int}
#     HybridSBPB200.set_metrics!(blk, crr, css, crs)                   # concatenated in vstarts layout
#     HybridSBPB200.set_bc!(blk, LFToB)                                # 4 x nelems Int matrix
#     HybridSBPB200.compute_tau!(blk, 2.0)                             # global_curved.jl:418-437
#     tr  = HybridSBPB200.Trace(blk, FToB, FToE, FToLF, EToO, EToS)    # connectivityarrays' outputs, 1-based
#     HybridSBPB200.local_setup!(blk; mode = HybridSBPB200.LOCAL_BAND)
#     HybridSBPB200.condense!(tr); HybridSBPB200.precond_setup!(tr); HybridSBPB200.coarse_setup!(tr; modes = 2)
#     λ, u, stats = HybridSBPB200.trace_solve(tr, g, gδ; tol = 1e-10)  # square_circle.jl:376-388
#
# or, leaving the reference's assembly and drivers untouched, only swap the plugin (square_circle.jl:297-299):
#
#     OPTYPE = typeof(HybridSBPB200.b200_factorization(ctx)(sparse([1], [1], [1.0])))
#     (M, FbarT, D, vstarts, FToλstarts) = LocalGlobalOperators(lop, Nr, Ns, FToB, FToE, FToLF, EToO, EToS,
#                                                               HybridSBPB200.b200_factorization(ctx))
#
# Lifetime.  The C objects keep raw pointers to their parents (trace -> blocks -> context).  Every wrapper therefore
# (a) holds a reference to its parent, so the parent is never collected first while the child is reachable, and
# (b) registers itself with the parent; `close(parent)` closes the children first, finalizers only call `close`, and
# `close` is idempotent -- the order in which the garbage collector runs finalizers no longer matters.
#
module HybridSBPB200

using LinearAlgebra
using SparseArrays
import LinearAlgebra: Factorization
import Base: \, size, close, adjoint

const libhsbp = get(ENV, "HSBP_LIB", joinpath(@__DIR__, "..", "hybridsbp_b200", "libhsbp.so"))

struct HsbpError <: Exception
  code::Cint
  msg::String
end

# children are kept as weak references: registering must not keep them alive
register!(parent, child) = (push!(parent.children, WeakRef(child)); child)
function close_children!(parent)
  for w in parent.children
    c = w.value
    c === nothing || close(c)
  end
  empty!(parent.children)
end

mutable struct Context
  h::Ptr{Cvoid}
  children::Vector{WeakRef}
  function Context(device::Integer = 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:hsbp_ctx_create, libhsbp), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h)
    rc == 0 || throw(HsbpError(rc, "hsbp_ctx_create failed (a B200 / sm_100 GPU is required; no CPU fallback)"))
    ctx = new(h[], WeakRef[])
    finalizer(close, ctx)
  end
end
"destroy the context after everything created on it (blocks, traces, factors, device vectors)"
function close(c::Context)
  c.h == C_NULL && return
  close_children!(c)
  ccall((:hsbp_ctx_destroy, libhsbp), Cint, (Ptr{Cvoid},), c.h)
  c.h = C_NULL
  nothing
end

lasterror(ctx::Context) = unsafe_string(ccall((:hsbp_last_error, libhsbp), Cstring, (Ptr{Cvoid},), ctx.h))
check(ctx::Context, rc) = rc == 0 ? nothing : throw(HsbpError(rc, lasterror(ctx)))

# ---- device vectors -------------------------------------------------------------------------------
mutable struct DeviceVector
  ctx::Context
  ptr::Ptr{Cvoid}
  n::Int
  function DeviceVector(ctx::Context, n::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:hsbp_malloc, libhsbp), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.h, 8n, p))
    v = register!(ctx, new(ctx, p[], n))
    finalizer(close, v)
  end
end
function close(v::DeviceVector)
  (v.ptr == C_NULL || v.ctx.h == C_NULL) && (v.ptr = C_NULL; return)
  ccall((:hsbp_free, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), v.ctx.h, v.ptr)
  v.ptr = C_NULL
  nothing
end
function DeviceVector(ctx::Context, a::AbstractVector{Float64})
  v = DeviceVector(ctx, length(a)); upload!(v, a); v
end
upload!(v::DeviceVector, a::AbstractVector{Float64}) =
  check(v.ctx, ccall((:hsbp_h2d, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Csize_t),
                     v.ctx.h, v.ptr, a, 8 * v.n))
function download(v::DeviceVector)
  a = Vector{Float64}(undef, v.n)
  check(v.ctx, ccall((:hsbp_d2h, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Csize_t),
                     v.ctx.h, a, v.ptr, 8 * v.n))
  a
end

# ---- block-local operators: replaces locoperator's assembled M̃ (global_curved.jl:211-506) ------------
mutable struct Blocks
  ctx::Context
  h::Ptr{Cvoid}
  p::Int
  Nr::Vector{Int64}
  Ns::Vector{Int64}
  vstarts::Vector{Int64}     # 1-based, as SBPLocalOperator1 builds it (global_curved.jl:685-686)
  children::Vector{WeakRef}
  function Blocks(ctx::Context, p::Integer, Nr::Vector{<:Integer}, Ns::Vector{<:Integer})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    nr, ns = Int64.(Nr), Int64.(Ns)
    check(ctx, ccall((:hsbp_blocks_create, libhsbp), Cint,
                     (Ptr{Cvoid}, Cint, Int64, Ptr{Int64}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                     ctx.h, p, length(nr), nr, ns, h))
    b = register!(ctx, new(ctx, h[], p, nr, ns, cumsum([1; (nr .+ 1) .* (ns .+ 1)]), WeakRef[]))
    finalizer(close, b)
  end
end
"destroy the blocks after the traces / BP1 stages built on them"
function close(b::Blocks)
  b.h == C_NULL && return
  close_children!(b)
  b.ctx.h == C_NULL || ccall((:hsbp_blocks_destroy, libhsbp), Cint, (Ptr{Cvoid},), b.h)
  b.h = C_NULL
  nothing
end
num_volume_points(b::Blocks) = ccall((:hsbp_blocks_num_volume_points, libhsbp), Int64, (Ptr{Cvoid},), b.h)
num_face_points(b::Blocks) = ccall((:hsbp_blocks_num_face_points, libhsbp), Int64, (Ptr{Cvoid},), b.h)

set_metrics!(b::Blocks, crr::Vector{Float64}, css::Vector{Float64}, crs::Vector{Float64}) =
  check(b.ctx, ccall((:hsbp_blocks_set_metrics, libhsbp), Cint,
                     (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), b.h, crr, css, crs))
set_bc!(b::Blocks, LFToB::AbstractMatrix{<:Integer}) =       # 4 x nelems, column-major = block by block
  check(b.ctx, ccall((:hsbp_blocks_set_bc, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Int64}), b.h, Int64.(vec(LFToB))))
compute_tau!(b::Blocks, tauscale::Real = 2.0) =
  check(b.ctx, ccall((:hsbp_blocks_compute_tau, libhsbp), Cint, (Ptr{Cvoid}, Cdouble), b.h, tauscale))

"y = M̃ u for every block (the SpMV lop[e].M̃ * u, global_curved.jl:470-492), device vectors"
apply!(y::DeviceVector, b::Blocks, u::DeviceVector) =
  check(b.ctx, ccall((:hsbp_apply, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), b.h, u.ptr, y.ptr))
"y = M̃ u and the per-block energies u_e' * (M̃ u)_e from the same pass (hsbp_apply_energy; line-marching kernel only)"
function apply_energy!(y::DeviceVector, b::Blocks, u::DeviceVector)
  en = Vector{Float64}(undef, length(b.Nr))
  check(b.ctx, ccall((:hsbp_apply_energy, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), b.h, u.ptr, y.ptr, en))
  en
end
"same through host arrays (H2D, kernels, D2H inside the call)"
function apply(b::Blocks, u::Vector{Float64})
  y = similar(u)
  check(b.ctx, ccall((:hsbp_apply_host, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), b.h, u, y))
  y
end

# ---- options, tau, face operators -----------------------------------------------------------------
"tuning / solver knobs of include/hsbp.h (\"fdm_gemm\", \"sweep_deep\", ...)"
set_option!(b::Blocks, name::AbstractString, value::Integer) =
  check(b.ctx, ccall((:hsbp_blocks_set_option, libhsbp), Cint, (Ptr{Cvoid}, Cstring, Int64), b.h, name, value))
"tau of faces 1..4 of every block, concatenated (the diagonals of lop[e].τ, global_curved.jl:418-437)"
function get_tau(b::Blocks)
  tau = Vector{Float64}(undef, num_face_points(b))
  check(b.ctx, ccall((:hsbp_blocks_get_tau, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.h, tau))
  tau
end
set_tau!(b::Blocks, tau::Vector{Float64}) =
  check(b.ctx, ccall((:hsbp_blocks_set_tau, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.h, tau))
"y += alpha * sum_k F_k v_k per block, block-face vector v (locbcarray!'s products, global_curved.jl:596-623)"
face_F_add!(y::DeviceVector, b::Blocks, v::DeviceVector, alpha::Real) =
  check(b.ctx, ccall((:hsbp_face_F_add, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), b.h, v.ptr, alpha, y.ptr))
"ft = F_k' u for faces 1..4 of every block (global_curved.jl:455-458)"
face_FT!(ft::DeviceVector, b::Blocks, u::DeviceVector) =
  check(b.ctx, ccall((:hsbp_face_FT, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), b.h, u.ptr, ft.ptr))
"tr = HfI_FT_k u (computetraction's operator, global_curved.jl:460-463, 638-644)"
face_traction!(tr::DeviceVector, b::Blocks, u::DeviceVector) =
  check(b.ctx, ccall((:hsbp_face_traction, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), b.h, u.ptr, tr.ptr))

# ---- the `factorization` plugin (global_curved.jl:659, 672, 698, 734) ------------------------------------
# SBPLocalOperator1 requires `factors[e] <: Factorization` and uses `F \ g`.  All blocks are solved in one
# batched call, so the per-block object is a view into a shared batch solver.
const LOCAL_PCG = 1
const LOCAL_CHOLESKY = 2
const LOCAL_BAND = 3      # banded Cholesky (blocks beyond the dense solver, e.g. the 201 x 201 block of BP1)
const LOCAL_FDM = 4       # PCG with the fast-diagonalisation preconditioner (uniform, large blocks)
struct LocalStats
  iterations_max::Int64
  iterations_sum::Int64
  failed_blocks::Int64
  max_rel_residual::Float64
end
local_setup!(b::Blocks; mode = LOCAL_PCG, tol = 1e-13, maxit = 100_000) =
  check(b.ctx, ccall((:hsbp_local_setup, libhsbp), Cint, (Ptr{Cvoid}, Cint, Cdouble, Int64), b.h, mode, tol, maxit))
function local_solve!(u::DeviceVector, b::Blocks, g::DeviceVector)
  st = Ref(LocalStats(0, 0, 0, 0.0))
  check(b.ctx, ccall((:hsbp_local_solve, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{LocalStats}),
                     b.h, g.ptr, u.ptr, st))
  st[]
end
"All-blocks solve as one Factorization: `F \\ g` with g the concatenated volume vector (matrix-free operator inside)."
struct B200LocalSolve <: Factorization{Float64}
  blocks::Blocks
end
size(F::B200LocalSolve) = (n = num_volume_points(F.blocks); (n, n))
function \(F::B200LocalSolve, g::AbstractVector{Float64})
  dg = DeviceVector(F.blocks.ctx, collect(g)); du = DeviceVector(F.blocks.ctx, length(g))
  st = local_solve!(du, F.blocks, dg)
  st.failed_blocks == 0 || error("local solve did not converge on $(st.failed_blocks) blocks: $(st)")
  download(du)
end

# The plugin itself, at the reference's seam: `factorization(x::SparseMatrixCSC)` -> an object `<: Factorization` with
# `F \ v` (global_curved.jl:734, square_circle.jl:383, odefun.jl:43) and `F' \ S` (global_curved.jl:774).
# SBPLocalOperator1 first calls the plugin on `sparse([1], [1], [1.0])` and takes `typeof` of the result as the element
# type of `factors` (global_curved.jl:681): the 1 x 1 probe goes through the same code path and yields the same type.
mutable struct B200Factorization <: Factorization{Float64}
  ctx::Context
  h::Ptr{Cvoid}
  n::Int
  function B200Factorization(ctx::Context, A::SparseMatrixCSC{Float64,<:Integer})
    n = size(A, 1)
    n == size(A, 2) || throw(DimensionMismatch("matrix is not square"))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ctx, ccall((:hsbp_factor_create, libhsbp), Cint,
                     (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
                     ctx.h, n, Int64.(A.colptr), Int64.(A.rowval), A.nzval, 1, h))      # 1-based arrays, untouched
    F = register!(ctx, new(ctx, h[], n))
    finalizer(close, F)
  end
end
B200Factorization(ctx::Context, A::Symmetric) = B200Factorization(ctx, sparse(A))
function close(F::B200Factorization)
  F.h == C_NULL && return
  F.ctx.h == C_NULL || ccall((:hsbp_factor_destroy, libhsbp), Cint, (Ptr{Cvoid},), F.h)
  F.h = C_NULL
  nothing
end
"`factorization = b200_factorization(ctx)` in place of `x -> cholesky(Symmetric(x))` (square_circle.jl:299, BP1.jl:78)"
b200_factorization(ctx::Context) = x -> B200Factorization(ctx, x)
size(F::B200Factorization) = (F.n, F.n)
adjoint(F::B200Factorization) = F                                   # symmetric: F' \ S is F \ S (global_curved.jl:774)
function \(F::B200Factorization, g::AbstractVector{<:Real})
  gin = Vector{Float64}(g)                                          # views (global_curved.jl:734) are copied out
  u = similar(gin)
  check(F.ctx, ccall((:hsbp_factor_solve, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), F.h, gin, u, 1))
  u
end
function \(F::B200Factorization, S::AbstractMatrix{<:Real})          # sparse or dense block of right-hand sides -> dense
  G = Matrix{Float64}(S)
  U = similar(G)
  check(F.ctx, ccall((:hsbp_factor_solve, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), F.h, G, U, size(G, 2)))
  U
end

# ---- trace (λ) operators and the Schur-complement solve (global_curved.jl:510-565, 730-797) --------------
mutable struct Trace
  blocks::Blocks
  h::Ptr{Cvoid}
  function Trace(b::Blocks, FToB::Vector{<:Integer}, FToE::AbstractMatrix{<:Integer}, FToLF::AbstractMatrix{<:Integer},
                 EToO::AbstractMatrix{Bool}, EToS::AbstractMatrix{<:Integer})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(b.ctx, ccall((:hsbp_trace_create, libhsbp), Cint,
                       (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt8}, Ptr{Int64}, Ref{Ptr{Cvoid}}),
                       b.h, length(FToB), Int64.(FToB), Int64.(vec(FToE)), Int64.(vec(FToLF)),
                       UInt8.(vec(EToO)), Int64.(vec(EToS)), h))
    t = register!(b, new(b, h[]))
    finalizer(close, t)
  end
end
function close(t::Trace)
  t.h == C_NULL && return
  (t.blocks.h == C_NULL || t.blocks.ctx.h == C_NULL) || ccall((:hsbp_trace_destroy, libhsbp), Cint, (Ptr{Cvoid},), t.h)
  t.h = C_NULL
  nothing
end
num_lambda(t::Trace) = ccall((:hsbp_trace_num_lambda, libhsbp), Int64, (Ptr{Cvoid},), t.h)
function FToλstarts(t::Trace, nfaces::Integer)
  s = Vector{Int64}(undef, nfaces + 1)
  check(t.blocks.ctx, ccall((:hsbp_trace_get_starts, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Int64}), t.h, s)); s
end
struct TraceStats                  # layout of hsbp_trace_stats
  outer_iterations::Int64
  converged::Int64
  rel_residual::Float64
  inner_iterations_sum::Int64
  inner_iterations_max::Int64
  local_solves::Int64
  true_rel_residual::Float64
  failed_local_blocks::Int64
  max_local_rel_residual::Float64
  coarse_dofs::Int64
  issued_iterations::Int64
  b_norm::Float64
  cg_loop_ms::Float64
end
TraceStats() = TraceStats(0, 0, 0.0, 0, 0, 0, 0.0, 0, 0.0, 0, 0, 0.0, 0.0)
"form the dense per-block S_e = F_eᵀ M̃_e⁻¹ F_e (assembleλmatrix's products, global_curved.jl:759-790); later solves use them"
condense!(t::Trace; enable::Bool = true) =
  check(t.blocks.ctx, ccall((:hsbp_trace_condense, libhsbp), Cint, (Ptr{Cvoid}, Cint), t.h, enable ? 1 : 0))

const PRECOND_JACOBI = 0
const PRECOND_FACE_BLOCKS = 1
"preconditioner of the CG on B: D, or the exact diagonal blocks B_ff (needs condense!)"
precond_setup!(t::Trace; kind::Integer = PRECOND_FACE_BLOCKS) =
  check(t.blocks.ctx, ccall((:hsbp_trace_precond_setup, libhsbp), Cint, (Ptr{Cvoid}, Cint), t.h, kind))
"second level: `modes` Legendre polynomials per face (0 = off); makes the CG iteration count independent of the number of blocks"
coarse_setup!(t::Trace; modes::Integer = 2) =
  check(t.blocks.ctx, ccall((:hsbp_trace_coarse_setup, libhsbp), Cint, (Ptr{Cvoid}, Cint), t.h, modes))
set_option!(t::Trace, name::AbstractString, value::Integer) =
  check(t.blocks.ctx, ccall((:hsbp_trace_set_option, libhsbp), Cint, (Ptr{Cvoid}, Cstring, Int64), t.h, name, value))
function last_local_stats(t::Trace)
  st = Ref(LocalStats(0, 0, 0, 0.0))
  check(t.blocks.ctx, ccall((:hsbp_trace_last_local_stats, libhsbp), Cint, (Ptr{Cvoid}, Ref{LocalStats}), t.h, st))
  st[]
end

# ---- multi-GPU: one Context per device = one NCCL rank; the library does every exchange (SURVEY.md section 8e) ----------
"128 bytes made by one rank; hand them to every other rank (MPI.jl, sockets, a file ...) and call comm_init! everywhere"
function comm_unique_id()
  id = Vector{UInt8}(undef, 128)
  rc = ccall((:hsbp_comm_unique_id, libhsbp), Cint, (Ptr{UInt8},), id)
  rc == 0 || throw(HsbpError(rc, "hsbp_comm_unique_id failed (NCCL not loadable?)"))
  id
end
comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, world::Integer) =
  check(ctx, ccall((:hsbp_comm_init, libhsbp), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), ctx.h, id, rank, world))
"""
    set_partition!(t, faces, partner, gamma, n_gamma_total)

After `Trace(...)` on the LOCAL connectivity of this rank's blocks (FToE = 0 for the side of a face that lives on another
rank): `faces[c]` 1-based local id of cut face c, `partner[c]` the rank holding its other side, `gamma[c]` its 0-based
index among all cut faces of the mesh.  Every rank calls it (also with no cut faces); afterwards condense!, precond_setup!,
coarse_setup! and trace_solve are collective.
"""
set_partition!(t::Trace, faces::Vector{<:Integer}, partner::Vector{<:Integer}, gamma::Vector{<:Integer}, n_gamma_total::Integer) =
  check(t.blocks.ctx, ccall((:hsbp_trace_set_partition, libhsbp), Cint, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int64),
                            t.h, length(faces), Int64.(faces), Int64.(partner), Int64.(gamma), n_gamma_total))

"how the last trace_solve exchanged data inside its iteration loop: 0 one rank, 1 NCCL, 2 peer memory over NVLink (option \"cg_p2p\")"
comm_path(t::Trace) = Int(ccall((:hsbp_trace_comm_path, libhsbp), Cint, (Ptr{Cvoid},), t.h))

# ---- SEAS BP1 ODE stage: replaces the body of odefun (seas/BP1/odefun.jl:8-121) ---------------------------
struct Bp1Params            # layout of hsbp_bp1_params
  Vp::Float64; mu_shear::Float64; sigma_n::Float64; eta::Float64; V0::Float64; tau_z0::Float64
  Dc::Float64; f0::Float64; b::Float64; ftol::Float64; atolx::Float64; rtolx::Float64; maxiter::Int64
end
struct Bp1Stats             # layout of hsbp_bp1_stats
  rejected::Int64; failure_bits::Int64; failed_nodes::Int64; newton_iterations_max::Int64; local_iterations::Int64
end
mutable struct Bp1Stage
  blocks::Blocks
  h::Ptr{Cvoid}
  function Bp1Stage(b::Blocks, RSa::Vector{Float64}, sJ::Vector{Float64}, prm::Bp1Params;
                    block::Integer = 1, fault_face::Integer = 1, loading_face::Integer = 2)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(b.ctx, ccall((:hsbp_bp1_create, libhsbp), Cint,
                       (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ref{Bp1Params}, Ref{Ptr{Cvoid}}),
                       b.h, block, fault_face, loading_face, RSa, sJ, Ref(prm), h))
    s = register!(b, new(b, h[]))
    finalizer(close, s)
  end
end
function close(s::Bp1Stage)
  s.h == C_NULL && return
  (s.blocks.h == C_NULL || s.blocks.ctx.h == C_NULL) || ccall((:hsbp_bp1_destroy, libhsbp), Cint, (Ptr{Cvoid},), s.h)
  s.h = C_NULL
  nothing
end
"condense the local solve onto the fault (N + 2 local solves once): every later odefun! is one small kernel"
condense!(s::Bp1Stage; enable::Bool = true) =
  check(s.blocks.ctx, ccall((:hsbp_bp1_condense, libhsbp), Cint, (Ptr{Cvoid}, Cint), s.h, enable ? 1 : 0))
"""
    odefun!(dψV, ψδ, p, t)   with p = (stage = Bp1Stage, reject_step = [false])

Drop-in for the reference's `odefun(dψV, ψδ, p, t)` (odefun.jl:8): boundary scatter, local solve, traction, per-node
root find and state evolution run on the device; a failure sets `p.reject_step[1] = true` exactly as odefun.jl:74-107 does,
so the `isoutofdomain = stepcheck` mechanism of BP1.jl:149-159 keeps working.
"""
function odefun!(dψV::Vector{Float64}, ψδ::Vector{Float64}, p, t)
  st = Ref(Bp1Stats(0, 0, 0, 0, 0))
  s = p.stage
  check(s.blocks.ctx, ccall((:hsbp_bp1_rhs, libhsbp), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Float64}, Ptr{Float64}, Ref{Bp1Stats}),
                            s.h, t, ψδ, dψV, st))
  st[].rejected != 0 && (p.reject_step[1] = true)
  nothing
end
"displacement field of the last odefun! call (u of odefun.jl:43)"
function displacement(s::Bp1Stage)
  u = Vector{Float64}(undef, num_volume_points(s.blocks))
  check(s.blocks.ctx, ccall((:hsbp_bp1_get_u, libhsbp), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, u))
  u
end

"λ = B⁻¹(gδ − F̄ᵀM̃⁻¹g), u = M̃⁻¹(g − F̄λ)   (square_circle.jl:376-388); returns (λ, u, stats)"
function trace_solve(t::Trace, g::Vector{Float64}, gδ::Vector{Float64}; tol = 1e-10, maxit = 10_000)
  ctx = t.blocks.ctx
  dg, dgd = DeviceVector(ctx, g), DeviceVector(ctx, gδ)
  dl, du = DeviceVector(ctx, length(gδ)), DeviceVector(ctx, length(g))
  st = Ref(TraceStats())
  check(ctx, ccall((:hsbp_trace_solve, libhsbp), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Ref{TraceStats}),
                   t.h, dg.ptr, dgd.ptr, dl.ptr, du.ptr, tol, maxit, st))
  download(dl), download(du), st[]
end

end # module
