# A mesh of straight-sided blocks -- configuration 2, meshes/flower_v2.inp (67 blocks, interfaces seen in reversed orientation from
# their plus side), or seas/BP1/meshes/BP1_v1.inp (194 blocks, two kinds of jump interfaces: side sets 7 and 8) -- solved with the
# REFERENCE's functions -- read_inp_2d, connectivityarrays, transfinite_blend (corner form), create_metrics, locoperator,
# LocalGlobalOperators, bcstarts, assembleλmatrix, locbcarray!, locsourcearray!, LocalToGLobalRHS!, computetraction -- which
# tests/refexec/minijulia.py takes from global_curved.jl under /root/reference.  The reference ships these meshes without drivers;
# this file is the test suite's own (inputs: the globals `meshfile`, `jumpcodes`, `slipshift`, `order` and `npts`).  The slip on the jump faces is a prescribed
# function of position rather than the jump of a manufactured solution, so that all three orientation branches of the jump data
# carry non-zero values.
include("global_curved.jl")

let
  mesh = read_inp_2d(meshfile)
  corners, blockcorner, blockface, facebc = mesh[1], mesh[2], mesh[3], mesh[4]
  nblocks = size(blockcorner, 2)
  nsides = length(facebc)
  conn = connectivityarrays(blockcorner, blockface)
  faceblock, facelocal, sameway, side = conn[1], conn[2], conn[3], conn[4]
  sizes = fill(npts, nblocks)

  # manufactured field (continuous across every interface) and the prescribed slip
  field(x, y) = sin.(0.9 .* x .+ 0.3) .* cos.(0.7 .* y .- 0.2) .+ 0.1 .* x .* y
  field_x(x, y) = 0.9 .* cos.(0.9 .* x .+ 0.3) .* cos.(0.7 .* y .- 0.2) .+ 0.1 .* y
  field_y(x, y) = -0.7 .* sin.(0.9 .* x .+ 0.3) .* sin.(0.7 .* y .- 0.2) .+ 0.1 .* x
  minus_laplacian(x, y) = (0.9^2 + 0.7^2) .* sin.(0.9 .* x .+ 0.3) .* cos.(0.7 .* y .- 0.2)
  slip(x, y) = 0.3 .* sin.(x .+ slipshift) .* cos.(2 .* y)      # slipshift = 0 on the flower mesh; the BP1 fault lies on x = 0

  # one straight-sided block per element: corner blend -> metrics -> local operator
  ops = Dict{Int64, Any}()
  for b = 1:nblocks
    cx = corners[1, blockcorner[:, b]]
    cy = corners[2, blockcorner[:, b]]
    mapx = (r, s) -> transfinite_blend(cx[1], cx[2], cx[3], cx[4], r, s)
    mapy = (r, s) -> transfinite_blend(cy[1], cy[2], cy[3], cy[4], r, s)
    ops[b] = locoperator(order, npts, npts, create_metrics(order, npts, npts, mapx, mapy), facebc[blockface[:, b]])
  end

  glob = LocalGlobalOperators(ops, sizes, sizes, facebc, faceblock, facelocal, sameway, side, A -> cholesky(Symmetric(A)))
  solvers, traceT, diagD, vstart, tstart = glob[1].F, glob[2], glob[3], glob[4], glob[5]
  jstart = bcstarts(facebc, faceblock, facelocal, jumpcodes, sizes, sizes)
  schur = assembleλmatrix(tstart, vstart, blockface, facebc, solvers, diagD, traceT)
  schur_solver = cholesky(Symmetric(schur))

  nvol = vstart[end] - 1
  ntrace = tstart[end] - 1
  jumps = zeros(jstart[end] - 1)
  for f = 1:nsides
    facebc[f] >= BC_JUMP_INTERFACE || continue
    b, lf = faceblock[1, f], facelocal[1, f]
    jumps[jstart[f]:(jstart[f+1]-1)] = slip(ops[b].facecoord[1][lf], ops[b].facecoord[2][lf])
  end

  # the face's slice of a trace-sized / jump-sized vector, in the order block b sees it on its local face lf
  seen_from(vec, starts, b, lf; aliased = false) = begin
    f = blockface[lf, b]
    rng = sameway[lf, b] ? (starts[f]:(starts[f+1]-1)) : ((starts[f+1]-1):-1:starts[f])
    aliased ? view(vec, rng) : vec[rng]
  end
  dirichlet = (lf, x, y, b) -> field(x, y)
  neumann = (lf, x, y, nx, ny, b) -> nx .* field_x(x, y) + ny .* field_y(x, y)
  half_of_this = (lf, x, y, b) -> (side[lf, b] == 1 ? -1 : 1) .* seen_from(jumps, jstart, b, lf)

  vol_rhs = zeros(nvol)
  trace_rhs = zeros(ntrace)
  for b = 1:nblocks
    mine = view(vol_rhs, vstart[b]:(vstart[b+1]-1))
    trace_parts = ntuple(lf -> seen_from(trace_rhs, tstart, b, lf; aliased = true), 4)
    locbcarray!(mine, trace_parts, ops[b], facebc[blockface[:, b]], dirichlet, neumann, half_of_this, (b))
    locsourcearray!(mine, (x, y) -> minus_laplacian(x, y), ops[b])
  end

  reduced = zeros(ntrace)
  work = zeros(nvol)
  LocalToGLobalRHS!(reduced, vol_rhs, trace_rhs, work, solvers, traceT, vstart)
  trace = schur_solver \ reduced
  coupled = vol_rhs - traceT' * trace
  solution = zeros(nvol)
  for b = 1:nblocks
    rows = vstart[b]:(vstart[b+1]-1)
    solution[rows] = solvers[b] \ coupled[rows]
  end

  fault_traction = zeros(length(jumps))
  for f = 1:nsides
    facebc[f] >= BC_JUMP_INTERFACE || continue
    b, lf = faceblock[1, f], facelocal[1, f]
    rows = vstart[b]:(vstart[b+1]-1)
    fault_traction[jstart[f]:(jstart[f+1]-1)] = computetraction(ops[b], lf, solution[rows], trace[tstart[f]:(tstart[f+1]-1)],
                                                                 jumps[jstart[f]:(jstart[f+1]-1)])
  end

  (verts = corners, EToV = blockcorner, EToF = blockface, FToB = facebc, FToE = faceblock, FToLF = facelocal, EToO = sameway, EToS = side,
   vstarts = vstart, FToλstarts = tstart, FToδstarts = jstart, FbarT = traceT, D = diagD, B = schur,
   δ = jumps, g = vol_rhs, gδ = trace_rhs, bλ = reduced, λ = trace, u = solution, τf = fault_traction)
end
