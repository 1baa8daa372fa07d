# Configuration 2 (meshes/flower_v2.inp: 67 blocks, interfaces with reversed orientation) driven through the REFERENCE's functions.
# The reference ships the mesh without a driver; this one is written for the tests, in the style of square_circle.jl, and is run
# by tests/refexec/minijulia.py with global_curved.jl included from /root/reference.  Unlike square_circle.jl the slip on the
# jump faces is a given function (not the jump of an exact solution), so that every branch of `in_jump` carries data.
include("global_curved.jl")

let
  (verts, EToV, EToF, FToB, EToDomain) = read_inp_2d("meshes/flower_v2.inp")
  (nelems, nfaces) = (size(EToV, 2), size(FToB, 1))
  (FToE, FToLF, EToO, EToS) = connectivityarrays(EToV, EToF)
  Nr = fill(N0, nelems)
  Ns = fill(N0, nelems)

  vex(x, y, e) = sin.(0.9 .* x .+ 0.3) .* cos.(0.7 .* y .- 0.2) .+ 0.1 .* x .* y
  vex_x(x, y, e) = 0.9 .* cos.(0.9 .* x .+ 0.3) .* cos.(0.7 .* y .- 0.2) .+ 0.1 .* y
  vex_y(x, y, e) = -0.7 .* sin.(0.9 .* x .+ 0.3) .* sin.(0.7 .* y .- 0.2) .+ 0.1 .* x
  laplace(x, y, e) = -(0.9^2 + 0.7^2) .* sin.(0.9 .* x .+ 0.3) .* cos.(0.7 .* y .- 0.2)
  slip(x, y) = 0.3 .* sin.(x) .* cos.(2 .* y)

  OPTYPE = typeof(locoperator(2, 16, 16))
  lop = Dict{Int64, OPTYPE}()
  for e = 1:nelems
    (x1, x2, x3, x4) = verts[1, EToV[:, e]]
    (y1, y2, y3, y4) = verts[2, EToV[:, e]]
    xt(r, s) = transfinite_blend(x1, x2, x3, x4, r, s)
    yt(r, s) = transfinite_blend(y1, y2, y3, y4, r, s)
    metrics = create_metrics(SBPp, Nr[e], Ns[e], xt, yt)
    lop[e] = locoperator(SBPp, Nr[e], Ns[e], metrics, FToB[EToF[:, e]])
  end

  (M, FbarT, D, vstarts, FToλstarts) = LocalGlobalOperators(lop, Nr, Ns, FToB, FToE, FToLF, EToO, EToS, (x) -> cholesky(Symmetric(x)))
  locfactors = M.F
  FToδstarts = bcstarts(FToB, FToE, FToLF, BC_JUMP_INTERFACE, Nr, Ns)
  VNp = vstarts[nelems+1]-1
  λNp = FToλstarts[nfaces+1]-1
  δNp = FToδstarts[nfaces+1]-1
  B = assembleλmatrix(FToλstarts, vstarts, EToF, FToB, locfactors, D, FbarT)
  BF = cholesky(Symmetric(B))

  (bλ, λ, gδ) = (zeros(λNp), zeros(λNp), zeros(λNp))
  (u, g) = (zeros(VNp), zeros(VNp))
  δ = zeros(δNp)
  for f = 1:nfaces
    if FToB[f] == BC_JUMP_INTERFACE
      e1 = FToE[1, f]
      lf1 = FToLF[1, f]
      (xf, yf) = lop[e1].facecoord
      δ[FToδstarts[f]:(FToδstarts[f+1]-1)] = slip(xf[lf1], yf[lf1])
    end
  end

  bc_Dirichlet = (lf, x, y, e, δ) -> vex(x, y, e)
  bc_Neumann   = (lf, x, y, nx, ny, e, δ) -> (nx .* vex_x(x, y, e) + ny .* vex_y(x, y, e))
  in_jump      = (lf, x, y, e, δ) -> begin
    f = EToF[lf, e]
    if EToS[lf, e] == 1
      return -δ[FToδstarts[f]:(FToδstarts[f+1]-1)]
    elseif EToO[lf, e]
      return  δ[FToδstarts[f]:(FToδstarts[f+1]-1)]
    else
      return  δ[(FToδstarts[f+1]-1):-1:FToδstarts[f]]
    end
  end

  for e = 1:nelems
    gδe = ntuple(4) do lf
      f = EToF[lf, e]
      if EToO[lf, e]
        return @view gδ[FToλstarts[f]:(FToλstarts[f+1]-1)]
      else
        return @view gδ[(FToλstarts[f+1]-1):-1:FToλstarts[f]]
      end
    end
    locbcarray!((@view g[vstarts[e]:vstarts[e+1]-1]), gδe, lop[e], FToB[EToF[:,e]], bc_Dirichlet, bc_Neumann, in_jump, (e, δ))
    source = (x, y, e) -> (-laplace(x, y, e))
    locsourcearray!((@view g[vstarts[e]:vstarts[e+1]-1]), source, lop[e], e)
  end

  LocalToGLobalRHS!(bλ, g, gδ, u, locfactors, FbarT, vstarts)
  λ[:] = BF \ bλ
  u[:] = -FbarT' * λ
  u[:] .= g .+ u
  for e = 1:nelems
    @views u[vstarts[e]:(vstarts[e+1]-1)] = locfactors[e] \ u[vstarts[e]:(vstarts[e+1]-1)]
  end

  # traction on the minus side of every jump face (computetraction, global_curved.jl:638-644)
  τf = zeros(δNp)
  for f = 1:nfaces
    if FToB[f] == BC_JUMP_INTERFACE
      e1 = FToE[1, f]
      lf1 = FToLF[1, f]
      λrng = FToλstarts[f]:(FToλstarts[f+1]-1)
      δrng = FToδstarts[f]:(FToδstarts[f+1]-1)
      urng = vstarts[e1]:(vstarts[e1+1]-1)
      τf[δrng] = computetraction(lop[e1], lf1, u[urng], λ[λrng], δ[δrng])
    end
  end

  (verts = verts, EToV = EToV, EToF = EToF, FToB = FToB, FToE = FToE, FToLF = FToLF, EToO = EToO, EToS = EToS,
   vstarts = vstarts, FToλstarts = FToλstarts, FToδstarts = FToδstarts, FbarT = FbarT, D = D, B = B,
   δ = δ, g = g, gδ = gδ, bλ = bλ, λ = λ, u = u, τf = τf)
end
