// Drives the C++ host mirror (block_b200/host) exactly as the reference's sweep drives its own classes, on one record of
// the real reference (tests/golden/*.npz re-written in the raw record format), and writes the results back for the Python
// test to compare.  usage: host_mirror_test <in.bin> <out.bin> [device]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "../../block_b200/host/b2d_host.hpp"

using namespace b2d_host;

struct Rec { std::vector<int> i; std::vector<double> d; };
typedef std::map<std::string, Rec> Records;

static Records read_records(const char* path) {
  Records out;
  std::ifstream f(path, std::ios::binary);
  while (true) {
    uint32_t n;
    if (!f.read((char*)&n, 4)) break;
    std::string name(n, ' ');
    f.read(&name[0], n);
    uint8_t dtype; f.read((char*)&dtype, 1);
    uint32_t nd; f.read((char*)&nd, 4);
    uint64_t cnt = 1;
    for (uint32_t k = 0; k < nd; ++k) { uint64_t d; f.read((char*)&d, 8); cnt *= d; }
    Rec r;
    if (dtype == 0) { std::vector<int32_t> t(cnt); f.read((char*)t.data(), 4 * cnt); r.i.assign(t.begin(), t.end()); }
    else { r.d.resize(cnt); f.read((char*)r.d.data(), 8 * cnt); }
    out[name] = r;
  }
  return out;
}
static void write_rec(std::ofstream& f, const std::string& name, const std::vector<double>& v) {
  uint32_t n = name.size(); f.write((char*)&n, 4); f.write(name.data(), n);
  uint8_t dt = 1; f.write((char*)&dt, 1);
  uint32_t nd = 1; f.write((char*)&nd, 4);
  uint64_t d = v.size(); f.write((char*)&d, 8);
  f.write((char*)v.data(), 8 * v.size());
}
static void write_rec(std::ofstream& f, const std::string& name, const std::vector<int>& v) {
  uint32_t n = name.size(); f.write((char*)&n, 4); f.write(name.data(), n);
  uint8_t dt = 0; f.write((char*)&dt, 1);
  uint32_t nd = 1; f.write((char*)&nd, 4);
  uint64_t d = v.size(); f.write((char*)&d, 8);
  std::vector<int32_t> t(v.begin(), v.end());
  f.write((char*)t.data(), 4 * t.size());
}

static int build_block(Records& rec, const std::string& p, SpinBlock& b) {
  const Rec& q = rec[p + "q"];
  const Rec& dims = rec[p + "dims"];
  for (size_t i = 0; i < dims.i.size(); ++i) {
    b.stateInfo.quanta.push_back(SpinQuantum(q.i[3 * i], q.i[3 * i + 1], q.i[3 * i + 2]));
    b.stateInfo.quantaStates.push_back(dims.i[i]);
  }
  b.sites = rec[p + "sites"].i;
  b.loopblock = rec[p + "flags"].i[0] != 0;
  int mask_mismatch = 0;
  const int nops = rec[p + "nops"].i[0];
  for (int m = 0; m < nops; ++m) {
    const std::string o = p + "op" + std::to_string(m) + ".";
    const std::vector<int>& meta = rec[o + "meta"].i;
    std::shared_ptr<SparseMatrix> op(new SparseMatrix());
    op->optype = (opTypes)meta[0];
    for (int k = 0; k < meta[2]; ++k) op->orbs.push_back(meta[3 + k]);
    op->comp = meta[5];
    op->deltaQuantum.assign(1, SpinQuantum(meta[6], meta[7], meta[8]));
    op->fermion = meta[9] != 0;
    op->allocate(b.stateInfo);                                  // the reference's allocation rule ...
    const std::vector<int>& allowed = rec[o + "allowed"].i;     // ... must reproduce the reference's mask bit for bit
    for (int i = 0; i < op->nrows(); ++i)
      for (int j = 0; j < op->ncols(); ++j)
        if ((op->allowed(i, j) != 0) != (allowed[(size_t)i * op->ncols() + j] != 0)) ++mask_mismatch;
    op->CollectFrom(rec[o + "data"].d);
    b.ops.push_back(op);
  }
  return mask_mismatch;
}

struct Functor : public Davidson_functor {
  const SpinBlock& big;
  explicit Functor(const SpinBlock& b) : big(b) {}
  void operator()(Wavefunction& c, Wavefunction& v) override { big.multiplyH(c, &v, 1); }   // davidson.C:19-22
  const SpinBlock& get_block() override { return big; }
};

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: host_mirror_test <in.bin> <out.bin> [device]\n"); return 2; }
  Records rec = read_records(argv[1]);
  std::ofstream out(argv[2], std::ios::binary);
  SpinBlock L, R, big;
  int mismatch = build_block(rec, "L.", L) + build_block(rec, "R.", R);
  write_rec(out, "mask_mismatch", std::vector<int>(1, mismatch));
  DeviceOptions opt;
  opt.device = argc > 3 ? atoi(argv[3]) : 0;
  opt.core_energy = rec["meta_f"].d[3];
  opt.ham = rec["meta"].i[7] == 1 ? HUBBARD : QUANTUM_CHEMISTRY;
  opt.norbs = (int)rec["spin_orbs_symmetry"].i.size() / 2;
  opt.deflation_min = rec["dav_in"].i[4];
  opt.deflation_max = rec["dav_in"].i[5];
  const std::vector<int>& tq = rec["psi_dq"].i;
  const SpinQuantum target(tq[0], tq[1], tq[2]);
  big.set_big_block(&L, &R, target, opt);

  // multiplyH: v += H c on a cleared v (linear.C:239-253)
  Wavefunction c(target, &big, false), v(target, &big, false);
  c.CollectFrom(rec["rpsi"].d);
  v.Clear();
  big.multiplyH(c, &v, 1);
  std::vector<double> flat;
  v.FlattenInto(flat);
  write_rec(out, "sigma", flat);

  DiagonalMatrix e;
  big.diagonalH(e);
  write_rec(out, "diag", e);

  // one TensorMultiply: (H_L x 1_R) c, the first term of multiplyH (spinblock.C:744)
  v.Clear();
  const SparseMatrix* ham = nullptr; const SparseMatrix* ovl = nullptr;
  for (auto& o : L.ops) if (o->optype == HAM) ham = o.get();
  for (auto& o : R.ops) if (o->optype == OVERLAP) ovl = o.get();
  operatorfunctions::TensorMultiply(&L, *ham, *ovl, &big, c, v, SpinQuantum(0, 0, 0), 1.0);
  v.FlattenInto(flat);
  write_rec(out, "tm_ham_left", flat);
  // and a transposed pair: Transposeview(c_i) on the right with CCDcomp_i on the left (opxop.C:276)
  const SparseMatrix* cre = nullptr; const SparseMatrix* ccd = nullptr;
  for (auto& o : R.ops) if (o->optype == CRE && !cre) cre = o.get();
  if (cre) for (auto& o : L.ops) if (o->optype == CRE_CRE_DESCOMP && o->orbs == cre->orbs) ccd = o.get();
  if (cre && ccd) {
    v.Clear();
    Transposeview top(*cre);
    operatorfunctions::TensorMultiply(&L, *ccd, top, &big, c, v, SpinQuantum(0, 0, 0), 1.0);
    v.FlattenInto(flat);
    write_rec(out, "tm_ccd_cre_t", flat);
  }

  const int nroots = rec["meta"].i[4];
  std::vector<Wavefunction> b(nroots), lower;
  for (int i = 0; i < nroots; ++i) { b[i].initialise(target, &big, false); b[i].CollectFrom(rec["guess" + std::to_string(i)].d); }
  DiagonalMatrix hd = rec["diag"].d;
  Functor fn(big);
  bool useprecond = true;
  Linear::block_davidson(b, hd, rec["dav_tol"].d[0], false, fn, useprecond, -1, lower);
  write_rec(out, "dav_evals", std::vector<double>(hd.begin(), hd.begin() + nroots));

  std::vector<Wavefunction> sol(nroots);
  for (int i = 0; i < nroots; ++i) { sol[i].initialise(target, &big, false); sol[i].CollectFrom(rec["guess" + std::to_string(i)].d); }
  std::vector<double> energies, spins;
  double error = 0.0;
  std::vector<Matrix> rotateMatrix;
  L.RenormaliseFrom(energies, spins, error, rotateMatrix, rec["meta"].i[5], 0, rec["dav_tol"].d[0], big, TRANSFORM, rec["rdm.args"].d[0], 0.0, false, L, L, R, true, false,
                    0, -1, lower, &sol, &rec["weights"].d);
  write_rec(out, "energies", energies);
  write_rec(out, "error", std::vector<double>(1, error));
  std::vector<int> kept;
  for (const Matrix& m : rotateMatrix) kept.push_back(m.Ncols());
  write_rec(out, "kept", kept);

  L.transform_operators(rotateMatrix);
  write_rec(out, "N.dims", L.stateInfo.quantaStates);
  L.ops[0]->FlattenInto(flat);
  write_rec(out, "N.op0.data", flat);
  L.ops.back()->FlattenInto(flat);
  write_rec(out, "N.oplast.data", flat);
  printf("host_mirror_test ok\n");
  return 0;
}
