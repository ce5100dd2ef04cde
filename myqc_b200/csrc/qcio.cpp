// File layer of PROGRAM int2e (src/integrals/int2e.f90:14-69) restated in C++:
//   getenv     src/myQC/env.f90:16-73      envdat / nucpos / fmem readers
//   buildBasis src/myQC/basis.f90:23-226   mybasis -> set / setinfo / bas / basinfo (+ text dumps)
//   Ftab       int2e.f90:161-163           Fortran unformatted sequential record
//   XX         int2e.f90:166,306-307       Fortran unformatted sequential record, gfortran subrecords
// and the program driver itself (myqc_int2e_main).
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/myqc_eri.h"
#include "../../include/myqc_int1e.h"

namespace myqc {
extern thread_local std::string g_last_error;
}
using myqc::g_last_error;

namespace {

int io_fail(const std::string& msg) {
    g_last_error = msg;
    return MYQC_ERR_IO;
}

std::string join(const char* dir, const char* name) {
    std::string d = (dir && *dir) ? dir : ".";
    if (d.back() != '/') d += '/';
    return d + name;
}

// list-directed token stream: whitespace/comma separated, `r*c` repeats expanded
bool read_tokens(const std::string& path, std::vector<std::string>& toks) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
        for (char& c : line)
            if (c == ',') c = ' ';
        std::istringstream ss(line);
        std::string t;
        while (ss >> t) {
            const size_t star = t.find('*');
            if (star != std::string::npos && star > 0) {
                const int rep = std::atoi(t.substr(0, star).c_str());
                for (int k = 0; k < rep; ++k) toks.push_back(t.substr(star + 1));
            } else {
                toks.push_back(t);
            }
        }
    }
    return true;
}

double to_double(std::string t) {
    for (char& c : t)
        if (c == 'D' || c == 'd') c = 'E';
    return std::strtod(t.c_str(), nullptr);
}

const char* kElements[10] = {"H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne"};  // basis.f90:68
const char* kBasisNames[4] = {"STO-3G", "tester1", "tester2", "tester3"};           // basis.f90:69

// gfortran list-directed dump of an INTEGER(4) array: 12-column fields, records wrapped at 80 columns
void write_int_list(FILE* f, const int32_t* v, size_t n) {
    size_t col = 0;
    for (size_t i = 0; i < n; ++i) {
        if (col + 12 > 80) { std::fputc('\n', f); col = 0; }
        std::fprintf(f, "%12d", v[i]);
        col += 12;
    }
    std::fputc('\n', f);
}

}  // namespace

extern "C" {

int myqc_read_env(const char* dir, int cap_nuc, int cap_opt, int* nnuc, int* nelcA, int* nelcB,
                  int32_t* atoms, double* xyz, double* fmem, int* nopt, int32_t* options) {
    std::vector<std::string> env, nuc, mem;
    if (!read_tokens(join(dir, "envdat"), env)) return io_fail("Couldn't open envdat");  // env.f90:37-41
    if (!read_tokens(join(dir, "nucpos"), nuc)) return io_fail("Couldn't open nucpos");  // :43-47
    if (!read_tokens(join(dir, "fmem"), mem)) return io_fail("Couldn't open fmem");      // :49-53
    if (mem.empty() || env.size() < 4) return io_fail("envdat/fmem truncated");
    if (fmem) *fmem = to_double(mem[0]);
    const int nn = std::atoi(env[0].c_str());
    const int na = std::atoi(env[1].c_str()), nb = std::atoi(env[2].c_str());
    const int no = std::atoi(env[3].c_str());
    if (nn < 1 || no < 0 || (int)env.size() < 4 + no) return io_fail("envdat malformed");
    if (nnuc) *nnuc = nn;
    if (nelcA) *nelcA = na;
    if (nelcB) *nelcB = nb;
    if (nopt) *nopt = no;
    if (options) {
        if (no > cap_opt) return io_fail("options capacity too small");
        for (int i = 0; i < no; ++i) options[i] = (int32_t)std::atoll(env[4 + i].c_str());
    }
    if (atoms || xyz) {
        if (nn > cap_nuc) return io_fail("nuclei capacity too small");
        if ((int)nuc.size() < 4 * nn) return io_fail("nucpos truncated");
        for (int i = 0; i < nn; ++i) {  // READ(2,*) atoms(i), xyz(i,0), xyz(i,1), xyz(i,2)  env.f90:65-67
            if (atoms) atoms[i] = (int32_t)std::atoi(nuc[4 * i].c_str());
            if (xyz)
                for (int c = 0; c < 3; ++c) xyz[i + (size_t)nn * c] = to_double(nuc[4 * i + 1 + c]);
        }
    }
    return MYQC_OK;
}

int myqc_build_basis(const char* mybasis_path, int bkey, int nnuc, const int32_t* atoms, int* nset_cap,
                     int* norb_cap, double* set, int32_t* setinfo, double* bas, int32_t* basinfo,
                     int* maxN, int* maxL, const char* out_dir) {
    if (!mybasis_path || !atoms || nnuc < 1 || bkey < 0 || bkey > 3) {
        g_last_error = "bad arguments to myqc_build_basis";
        return MYQC_ERR_BAD_ARG;
    }
    // rows of whitespace-separated tokens, blank lines dropped (list-directed READ skips them)
    std::ifstream f(mybasis_path);
    if (!f) return io_fail("Couldn't open mybasis");
    std::vector<std::vector<std::string>> rows;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::vector<std::string> r;
        std::string t;
        while (ss >> t) r.push_back(t);
        if (!r.empty()) rows.push_back(r);
    }
    size_t start = 0;
    for (; start < rows.size(); ++start)
        if (rows[start][0] == kBasisNames[bkey]) break;  // basis.f90:95-97
    if (start + 2 >= rows.size() || rows[start + 1].size() < 5 || rows[start + 2].size() < 2)
        return io_fail("basis set not found in mybasis");
    const int Omax = std::atoi(rows[start + 1][2].c_str());
    const int almax = std::atoi(rows[start + 1][3].c_str());
    const int OpS = std::atoi(rows[start + 1][4].c_str());  // :101
    // Callers size set/bas/setinfo for 4 orbitals per set (s or s,px,py,pz: all the ERI engine evaluates, and what the
    // reference's mybasis holds); any other header value is refused here instead of overrunning their buffers.
    if (OpS != 4) {
        g_last_error = "mybasis: orbitals per set must be 4 (s / sp sets), found " + std::to_string(OpS);
        return MYQC_ERR_UNSUPPORTED;
    }
    const int mN = std::atoi(rows[start + 2][0].c_str()), mL = std::atoi(rows[start + 2][1].c_str());  // :102
    if (maxN) *maxN = mN;
    if (maxL) *maxL = mL;
    const int setl = 3 + OpS;
    const size_t n_bas = (size_t)nnuc * almax * OpS, n_basinfo = 2 + (size_t)5 * Omax * nnuc;
    const size_t n_set = (size_t)nnuc * almax, n_setinfo = 2 + (size_t)nnuc * almax * setl;
    // sizing call
    if (!set || !setinfo || !bas || !basinfo) {
        if (nset_cap) *nset_cap = (int)n_set;
        if (norb_cap) *norb_cap = Omax * nnuc;
        return MYQC_OK;
    }
    if ((nset_cap && *nset_cap < (int)n_set) || (norb_cap && *norb_cap < Omax * nnuc)) {
        g_last_error = "output capacity too small";
        return MYQC_ERR_BAD_ARG;
    }
    std::fill(bas, bas + n_bas, 0.0);
    std::fill(set, set + n_set, 0.0);
    std::fill(basinfo, basinfo + n_basinfo, 0);
    std::fill(setinfo, setinfo + n_setinfo, 0);
    int setnum = 0, orbnum = 0;
    for (int i = 0; i < nnuc; ++i) {
        if (atoms[i] < 1 || atoms[i] > 10) return io_fail("element not in mybasis (H..Ne only)");
        const char* sym = kElements[atoms[i] - 1];
        size_t r = start + 1;
        for (; r < rows.size(); ++r)
            if (rows[r][0] == sym) break;  // :116-119
        if (r + 1 >= rows.size() || rows[r + 1].size() < 3) return io_fail("atom not found in mybasis");
        const int sec = std::atoi(rows[r + 1][0].c_str()), orb = std::atoi(rows[r + 1][1].c_str()),
                  nset = std::atoi(rows[r + 1][2].c_str());  // :125
        r += 2;
        basinfo[0] = OpS;
        basinfo[1] += orb;
        setinfo[0] += nset;
        setinfo[1] = setl;
        for (int j = 0; j < sec; ++j) {
            if (r >= rows.size() || rows[r].size() < 5) return io_fail("mybasis section header malformed");
            const int func = std::atoi(rows[r][0].c_str()), coef = std::atoi(rows[r][1].c_str()),
                      pri = std::atoi(rows[r][2].c_str()), ang = std::atoi(rows[r][3].c_str()),
                      ori = std::atoi(rows[r][4].c_str());  // :136
            ++r;
            for (int k = 0; k < func; ++k, ++r) {
                if (r >= rows.size() || (int)rows[r].size() < coef + 1) return io_fail("mybasis primitive line malformed");
                std::vector<double> val(coef);
                for (int c = 0; c < coef; ++c) val[c] = to_double(rows[r][c]);
                const double temp = to_double(rows[r][coef]);  // :143
                const int setn = (int)std::lround(temp);       // NINT, :146
                const int s = setnum + setn;
                if (s < 0 || (size_t)s >= n_set) return io_fail("set id out of range in mybasis");
                set[s] = val[coef - 1];                        // :147
                int setorbs = setinfo[1 + s * setl + 1];       // :148
                setinfo[1 + s * setl + 3] = i;                 // :149
                if (ori == -1) {                               // :154-157
                    if (setorbs + 1 > OpS) return io_fail("too many orbitals in a set");
                    setinfo[1 + s * setl + 4 + setorbs] = orbnum;
                    bas[setorbs + (size_t)s * OpS] = val[0];
                    ++setorbs;
                } else if (ori == 2) {                         // :160-167
                    if (setorbs + 3 > OpS) return io_fail("too many orbitals in a set");
                    for (int m = 0; m < 3; ++m) {
                        setinfo[1 + s * setl + 4 + setorbs] = orbnum + m;
                        bas[setorbs + (size_t)s * OpS] = val[0];
                        ++setorbs;
                    }
                    if (setinfo[1 + s * setl + 2] < 1) setinfo[1 + s * setl + 2] = 1;
                } else {
                    g_last_error = "that angular quantum number not implemented (basis.f90:170-174)";
                    return MYQC_ERR_UNSUPPORTED;
                }
                setinfo[1 + s * setl + 1] = setorbs;           // :176
            }
            if (ori == -1) {                                   // :182-184
                const int32_t v[5] = {pri, ang, ori, func, i};
                std::memcpy(&basinfo[2 + 5 * orbnum], v, sizeof(v));
                orbnum += 1;
            } else {                                           // :187-191
                for (int m = 0; m < 3; ++m) {
                    const int32_t v[5] = {pri, ang, m, func, i};
                    std::memcpy(&basinfo[2 + 5 * (orbnum + m)], v, sizeof(v));
                }
                orbnum += 3;
            }
        }
        setnum = setinfo[0];  // :205
    }
    if (out_dir) {  // :213-218
        FILE* fb = std::fopen(join(out_dir, "basinfo").c_str(), "w");
        FILE* fs = std::fopen(join(out_dir, "setinfo").c_str(), "w");
        if (!fb || !fs) {
            if (fb) std::fclose(fb);
            if (fs) std::fclose(fs);
            return io_fail("cannot write basinfo/setinfo");
        }
        write_int_list(fb, basinfo, n_basinfo);
        std::fprintf(fb, "\n #orbitals per set, #orbitals, {principle qn., angular qn., orientation, #primatives, center number},... \n");
        write_int_list(fs, setinfo, n_setinfo);
        std::fprintf(fs, "\n #sets, length of each set,{#orbitals, max ang qn., center num, [orbital 0, orbital 1, ...]},...\n");
        std::fclose(fb);
        std::fclose(fs);
    }
    return MYQC_OK;
}

int myqc_read_ftab(const char* path, double* ftab) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return io_fail("Couldn't open Ftab");
    int32_t m0 = 0, m1 = 0;
    const size_t n = 121 * 23;
    bool ok = std::fread(&m0, 4, 1, f) == 1 && m0 == (int32_t)(n * 8) && std::fread(ftab, 8, n, f) == n &&
              std::fread(&m1, 4, 1, f) == 1 && m1 == m0;
    std::fclose(f);
    if (!ok) return io_fail("Ftab is not a 121x23 float64 Fortran record");
    return MYQC_OK;
}

// gfortran's default maximum subrecord payload (libgfortran io/unix.h: MAX_RECORD 0x7ffffff7)
static const int64_t kMaxSub = 2147483639;

int myqc_write_xx(const char* path, const double* xx, int norb) {
    return myqc_write_xx_ex(path, xx, norb, kMaxSub);
}

int myqc_write_xx_ex(const char* path, const double* xx, int norb, int64_t max_subrecord) {
    if (max_subrecord < 8 || max_subrecord > kMaxSub) max_subrecord = kMaxSub;
    FILE* f = std::fopen(path, "wb");
    if (!f) return io_fail("cannot open XX for writing");
    const int64_t n = norb;
    int64_t left = n * n * n * n * 8;
    const char* p = reinterpret_cast<const char*>(xx);
    bool first = true, ok = true;
    do {
        const int64_t chunk = left > max_subrecord ? max_subrecord : left;
        const bool more = left > chunk;
        // leading marker negative: continued in the next subrecord; trailing negative: has a predecessor
        const int32_t lead = (int32_t)(more ? -chunk : chunk);
        const int32_t trail = (int32_t)(first ? chunk : -chunk);
        ok = ok && std::fwrite(&lead, 4, 1, f) == 1;
        ok = ok && (chunk == 0 || std::fwrite(p, 1, (size_t)chunk, f) == (size_t)chunk);
        ok = ok && std::fwrite(&trail, 4, 1, f) == 1;
        p += chunk;
        left -= chunk;
        first = false;
    } while (left > 0 && ok);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) return io_fail("short write on XX");
    return MYQC_OK;
}

int myqc_read_xx(const char* path, double* xx, int norb) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return io_fail("cannot open XX");
    const int64_t n = norb;
    int64_t left = n * n * n * n * 8;
    char* p = reinterpret_cast<char*>(xx);
    bool ok = true, more = true;
    while (ok && more) {
        int32_t lead = 0, trail = 0;
        ok = std::fread(&lead, 4, 1, f) == 1;
        const int64_t chunk = lead < 0 ? -(int64_t)lead : lead;
        more = lead < 0;
        ok = ok && chunk <= left && (chunk == 0 || std::fread(p, 1, (size_t)chunk, f) == (size_t)chunk);
        ok = ok && std::fread(&trail, 4, 1, f) == 1 && (trail == chunk || trail == -chunk);
        p += chunk;
        left -= chunk;
    }
    std::fclose(f);
    if (!ok || left != 0) return io_fail("XX record does not hold norb^4 doubles");
    return MYQC_OK;
}

static void touch_error(const char* dir) {  // CALL EXECUTE_COMMAND_LINE('touch error')
    FILE* f = std::fopen(join(dir, "error").c_str(), "a");
    if (f) std::fclose(f);
}
static bool exists(const std::string& p) {
    FILE* f = std::fopen(p.c_str(), "rb");
    if (f) std::fclose(f);
    return f != nullptr;
}
static void write_fmem(const char* dir, double fmem) {  // nmem / setenv, env.f90:88-96,106-112
    FILE* f = std::fopen(join(dir, "fmem").c_str(), "w");
    if (!f) return;
    std::fprintf(f, "%26.16f     \n", fmem);
    std::fclose(f);
}

static int int2e_main_impl(const char* dir, int ngpu);
int myqc_int2e_main(const char* dir, int ngpu) {
    const int rc = int2e_main_impl(dir, ngpu);
    std::fflush(stdout);  // see myqc_parse_main
    return rc;
}
static int int2e_main_impl(const char* dir, int ngpu) {
    std::printf("\n int2e called\n");  // int2e.f90:48-49
    int nnuc = 0, nA = 0, nB = 0, nopt = 0;
    double fmem = 0;
    int rc = myqc_read_env(dir, 0, 0, &nnuc, &nA, &nB, nullptr, nullptr, &fmem, &nopt, nullptr);
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    std::vector<int32_t> atoms(nnuc), options(nopt > 17 ? nopt : 17, 0);
    std::vector<double> xyz((size_t)3 * nnuc);
    rc = myqc_read_env(dir, nnuc, (int)options.size(), &nnuc, &nA, &nB, atoms.data(), xyz.data(), &fmem, &nopt, options.data());
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    if (exists(join(dir, "error"))) return MYQC_OK;  // :51-52 (STOP)

    // buildBasis (:55) -- also rewrites basinfo / setinfo like the reference
    int nset_cap = 0, norb_cap = 0, maxN = 0, maxL = 0;
    const std::string mybasis = join(dir, "mybasis");
    rc = myqc_build_basis(mybasis.c_str(), options[2], nnuc, atoms.data(), &nset_cap, &norb_cap, nullptr, nullptr, nullptr, nullptr, &maxN, &maxL, nullptr);
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    const int OpS = 4, setl = 7;
    std::vector<double> set(nset_cap), bas((size_t)nset_cap * OpS);
    std::vector<int32_t> setinfo(2 + (size_t)nset_cap * setl), basinfo(2 + (size_t)5 * norb_cap);
    rc = myqc_build_basis(mybasis.c_str(), options[2], nnuc, atoms.data(), &nset_cap, &norb_cap, set.data(), setinfo.data(), bas.data(), basinfo.data(), &maxN, &maxL, dir);
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }

    if (exists(join(dir, "XX"))) {  // :58-63
        std::printf(" Reading two electron integrals from intermediate\n");
    } else {
        std::printf(" Constructing two electron integrals\n");
        // proc2e :78-351
        const int nset = setinfo[0], norb = basinfo[1];
        const int setK = OpS * OpS * (2 * maxL * (2 * maxL + 1) / 2) * (2 * maxL * (2 * maxL + 1) / 2) * (2 * maxL * (2 * maxL + 1) / 2);
        long long npri = 0;
        for (int i = 0; i < norb; ++i) npri += basinfo[1 + i * 5 + 4];
        std::printf(" Total primative integrals : %20lld\n", npri * npri * npri * npri);
        std::printf(" Total to be evaluated : %20lld\n", ((npri + 1) * npri) * ((npri + 1) * npri) / 8 + ((npri + 1) * npri) / 4);
        std::vector<double> ft(121 * 23);
        rc = myqc_read_ftab(join(dir, "Ftab").c_str(), ft.data());
        if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
        const double n4 = (double)norb * norb * norb * norb;
        const double temp = n4 * 8.0 / 1.0e6 + (double)setK * norb * norb * 8.0 / 1.0e6;  // :169-170
        std::printf(" Allocating space for intermediate matrices (MB)  %8.5f\n", temp);
        double* xx = static_cast<double*>(std::malloc((size_t)n4 * sizeof(double)));
        if (!xx) {  // :174-178
            std::printf(" Could not allocate memory in int2e:proc2e\n");
            touch_error(dir);
            return MYQC_ERR_NOMEM;
        }
        fmem -= temp;
        write_fmem(dir, fmem);
        rc = myqc_eri_dense(nnuc, xyz.data(), nset, setl, set.data(), setinfo.data(), OpS, bas.data(), basinfo.data(), ft.data(), xx, ngpu);
        if (!rc) rc = myqc_write_xx(join(dir, "XX").c_str(), xx, norb);
        if (rc) {
            std::printf(" int2e: %s\n", myqc_last_error());
            std::remove(join(dir, "XX").c_str());
            touch_error(dir);
            std::free(xx);
            return rc;
        }
        if (options[7] >= 3) {  // :310-331
            std::printf(" Two electron Integrals by Orbital\n");
            std::printf("    I   J   G   H    Value\n");
            const long long n = norb;
            auto X = [&](long long i, long long j, long long g, long long h) { return xx[i + n * (j + n * (g + n * h))]; };
            for (int i = 0; i < norb; ++i)
                for (int j = i; j < norb; ++j) {
                    for (int h = j; h < norb; ++h)
                        if (X(i, j, i, h) != 0.0) std::printf("    %3d %3d %3d %3d    %15.8f\n", i, j, i, h, X(i, j, i, h));
                    for (int g = i + 1; g < norb; ++g)
                        for (int h = g; h < norb; ++h)
                            if (X(i, j, g, h) != 0.0) std::printf("    %3d %3d %3d %3d    %15.8f\n", i, j, g, h, X(i, j, g, h));
                }
        }
        std::free(xx);
        fmem += temp;  // :343-345
        write_fmem(dir, fmem);
        std::printf("\n Two electron integrals constructed on the GPU\n");
    }
    write_fmem(dir, fmem);  // setenv :69
    return MYQC_OK;
}

// PROGRAM int1e, src/integrals/int1e.f90:14-131 --------------------------------------------------
// list-directed dump of a REAL(8) array: three values per line, 17 significant digits (what the
// consumers read back with READ(u,*), scf.f90:140-144)
static int write_real_list(const std::string& path, const double* v, size_t n) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return io_fail("cannot write " + path);
    for (size_t i = 0; i < n; ++i) {
        std::fprintf(f, "  %24.16E", v[i]);
        if (i % 3 == 2 || i + 1 == n) std::fputc('\n', f);
    }
    std::fclose(f);
    return MYQC_OK;
}

static int int1e_main_impl(const char* dir);
int myqc_int1e_main(const char* dir) {
    const int rc = int1e_main_impl(dir);
    std::fflush(stdout);
    return rc;
}
static int int1e_main_impl(const char* dir) {
    std::printf(" int1e called\n");  // int1e.f90:47
    int nnuc = 0, nA = 0, nB = 0, nopt = 0;
    double fmem = 0;
    int rc = myqc_read_env(dir, 0, 0, &nnuc, &nA, &nB, nullptr, nullptr, &fmem, &nopt, nullptr);
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    std::vector<int32_t> atoms(nnuc), options(nopt > 17 ? nopt : 17, 0);
    std::vector<double> xyz((size_t)3 * nnuc);
    rc = myqc_read_env(dir, nnuc, (int)options.size(), &nnuc, &nA, &nB, atoms.data(), xyz.data(), &fmem, &nopt, options.data());
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    if (exists(join(dir, "error"))) return MYQC_OK;  // :49-50 (STOP)

    // buildBasis (:58) -- writes basinfo / setinfo, which every later stage reads for norb
    int nset_cap = 0, norb_cap = 0, maxN = 0, maxL = 0;
    const std::string mybasis = join(dir, "mybasis");
    rc = myqc_build_basis(mybasis.c_str(), options[2], nnuc, atoms.data(), &nset_cap, &norb_cap, nullptr, nullptr, nullptr, nullptr, &maxN, &maxL, nullptr);
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    const int OpS = 4, setl = 7;
    std::vector<double> set(nset_cap), bas((size_t)nset_cap * OpS);
    std::vector<int32_t> setinfo(2 + (size_t)nset_cap * setl), basinfo(2 + (size_t)5 * norb_cap);
    rc = myqc_build_basis(mybasis.c_str(), options[2], nnuc, atoms.data(), &nset_cap, &norb_cap, set.data(), setinfo.data(), bas.data(), basinfo.data(), &maxN, &maxL, dir);
    if (rc) { std::printf(" %s\n", myqc_last_error()); touch_error(dir); return rc; }
    const int nset = setinfo[0], norb = basinfo[1];
    long long npri = 0;
    for (int i = 0; i < norb; ++i) npri += basinfo[1 + i * 5 + 4];
    std::printf("\n Number of orbitals     %d\n Number of primatives   %lld\n\n", norb, npri);  // :70-75
    const double temp = 2.0 * norb * norb * 8.0 / 1.0e6;  // :77-78
    std::printf(" Allocating space for int1e (MB) %8.5f\n", temp);
    if (fmem - temp < 0.0) {  // :79-82
        std::printf(" int1e: max memory reached\n");
        touch_error(dir);
        return MYQC_ERR_NOMEM;
    }
    if (exists(join(dir, "Suv")) && exists(join(dir, "Huv"))) {  // :98-111
        std::fclose(std::fopen(join(dir, "Sold").c_str(), "a"));
        std::fclose(std::fopen(join(dir, "Hold").c_str(), "a"));
        std::printf(" Reading overlap matrix from Suv\n Reading Fock matrix from Huv\n");
    } else {
        std::vector<double> ft(121 * 23), S((size_t)norb * norb), H((size_t)norb * norb);
        rc = myqc_read_ftab(join(dir, "Ftab").c_str(), ft.data());
        if (!rc) rc = myqc_int1e(nnuc, xyz.data(), atoms.data(), nset, setl, set.data(), setinfo.data(), OpS, bas.data(), basinfo.data(), ft.data(), S.data(), H.data());
        if (!rc) rc = write_real_list(join(dir, "Suv"), S.data(), S.size());
        if (!rc) rc = write_real_list(join(dir, "Huv"), H.data(), H.size());
        if (rc) { std::printf(" int1e: %s\n", myqc_last_error()); touch_error(dir); return rc; }
        std::printf(" Overlap written to Suv\n One electron integrals written to Huv\n");
    }
    write_fmem(dir, fmem);  // the ledger is debited and credited by the same amount (:78,126), setenv :127
    return MYQC_OK;
}

}  // extern "C"
