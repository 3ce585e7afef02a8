/* h5_check.c -- a second, independent reader of the HDF5 files hdf5.py writes (test infrastructure; shares no code with the Python writer /
 * reader).  libhdf5, h5py and HDF5.jl do not exist in this image, so this program checks the files against the HDF5 File Format
 * Specification (version 1.8, the structures every libhdf5 reads) byte by byte and prints a canonical listing the test compares with what
 * was written:
 *
 *   superblock v0 (III.A): signature, version numbers, size of offsets / lengths = 8, group leaf / internal K, base address 0, end-of-file
 *       address = file size, root symbol-table entry;
 *   groups: object header v1 (IV.A.1.a) with a Symbol Table message 0x0011 -> v1 B-tree "TREE" node type 0 (III.A.1) -> symbol-table nodes
 *       "SNOD" (III.B) whose entries name their links through the local heap "HEAP" (III.D);
 *   datasets: Dataspace v1 (0x0001), Datatype (0x0003: class 0 fixed point, 1 floating point, 6 compound), Data Layout v3 contiguous
 *       (0x0008); header sizes, 8-byte message alignment and address ranges are verified on the way.
 *
 * Usage: h5_check file.h5   ->  one line per object:  <path> <class><bits> [dims] first values...   exit code 0 = every check passed. */
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static unsigned char* F;
static uint64_t FSZ;
static int failures = 0;

#define CHECK(cond, ...) do { if (!(cond)) { fprintf(stderr, "CHECK FAILED: " __VA_ARGS__); fprintf(stderr, "\n"); failures++; } } while (0)

static uint64_t rd(uint64_t off, int n) {
    uint64_t v = 0;
    if (off + n > FSZ) { fprintf(stderr, "read past end of file at %" PRIu64 "\n", off); exit(2); }
    for (int i = n - 1; i >= 0; i--) v = (v << 8) | F[off + i];
    return v;
}

static void walk_group(uint64_t btree, uint64_t heap, const char* prefix, int leaf_k);

/* datatype message at `off`: prints class + size, returns element size; *is_complex set for the {r, i} compound */
static uint32_t datatype(uint64_t off, char* out, size_t outn) {
    const unsigned cls = F[off] & 0x0F, ver = F[off] >> 4;
    const uint32_t size = (uint32_t)rd(off + 4, 4);
    CHECK(ver == 1, "datatype version %u (expected 1)", ver);
    if (cls == 0) {                                  /* fixed point: bit 3 of class bits 0 = signed */
        CHECK((F[off + 1] & 1) == 0, "fixed-point byte order must be little endian");
        CHECK(rd(off + 8, 2) == 0 && rd(off + 10, 2) == 8u * size, "fixed-point bit offset / precision");
        snprintf(out, outn, "%s%u", (F[off + 1] & 0x08) ? "int" : "uint", 8 * size);
    } else if (cls == 1) {                           /* IEEE floating point */
        CHECK((F[off + 1] & 1) == 0, "floating-point byte order must be little endian");
        CHECK(rd(off + 10, 2) == 8u * size, "floating-point precision");
        if (size == 8) CHECK(F[off + 12] == 52 && F[off + 13] == 11 && F[off + 15] == 52 && rd(off + 16, 4) == 1023 && F[off + 2] == 63, "IEEE double fields");
        if (size == 4) CHECK(F[off + 12] == 23 && F[off + 13] == 8 && F[off + 15] == 23 && rd(off + 16, 4) == 127 && F[off + 2] == 31, "IEEE single fields");
        snprintf(out, outn, "float%u", 8 * size);
    } else if (cls == 6) {                           /* compound: HDF5.jl's ComplexF64 = {r: f64 @ 0, i: f64 @ 8} */
        const unsigned nmemb = (unsigned)rd(off + 1, 2);
        CHECK(nmemb == 2 && size == 16, "compound must be the 16-byte {r, i} pair (members %u, size %u)", nmemb, size);
        uint64_t p = off + 8;
        const char* names[2] = {"r", "i"};
        for (unsigned m = 0; m < nmemb && m < 2; m++) {
            CHECK(strcmp((const char*)F + p, names[m]) == 0, "compound member %u is named '%s'", m, (const char*)F + p);
            p += (strlen((const char*)F + p) + 8) / 8 * 8;
            CHECK(rd(p, 4) == 8u * m, "compound member byte offset");
            CHECK(F[p + 4] == 0, "compound member dimensionality must be 0");
            p += 4 + 1 + 3 + 4 + 4 + 16;             /* offset, dimensionality, reserved, permutation, reserved, 4 dimension sizes */
            char tmp[32];
            CHECK(datatype(p, tmp, sizeof tmp) == 8 && strcmp(tmp, "float64") == 0, "compound member type must be float64");
            p += 8 + 12;
        }
        snprintf(out, outn, "complex128");
    } else {
        CHECK(0, "unexpected datatype class %u", cls);
        snprintf(out, outn, "class%u", cls);
    }
    return size;
}

static void object(uint64_t addr, const char* path, int leaf_k) {
    CHECK(addr % 8 == 0 && addr + 16 <= FSZ, "object header address %" PRIu64, addr);
    CHECK(F[addr] == 1, "object header version %u at %" PRIu64, F[addr], addr);
    const unsigned nmsg = (unsigned)rd(addr + 2, 2);
    const uint32_t hsize = (uint32_t)rd(addr + 8, 4);
    CHECK(rd(addr + 4, 4) >= 1, "object reference count");
    uint64_t p = addr + 16;                          /* 12 bytes of prefix fields + 4 bytes alignment */
    const uint64_t end = p + hsize;
    CHECK(end <= FSZ, "object header runs past the end of the file");
    int have_space = 0, have_type = 0, have_layout = 0, ndim = 0;
    uint64_t dims[8] = {0}, daddr = 0, dsize = 0, stab_btree = 0, stab_heap = 0;
    char tname[32] = "?";
    uint32_t esize = 0;
    for (unsigned m = 0; m < nmsg; m++) {
        CHECK(p + 8 <= end, "message %u of %s starts outside the header", m, path);
        const unsigned type = (unsigned)rd(p, 2), size = (unsigned)rd(p + 2, 2);
        CHECK(size % 8 == 0, "message size %u is not a multiple of 8", size);
        const uint64_t b = p + 8;
        if (type == 0x0001) {                        /* dataspace v1 */
            CHECK(F[b] == 1, "dataspace version %u", F[b]);
            ndim = F[b + 1];
            CHECK(ndim <= 8 && F[b + 2] == 0, "dataspace rank %d / flags %u", ndim, F[b + 2]);
            for (int k = 0; k < ndim; k++) dims[k] = rd(b + 8 + 8 * k, 8);
            have_space = 1;
        } else if (type == 0x0003) {
            esize = datatype(b, tname, sizeof tname);
            have_type = 1;
        } else if (type == 0x0008) {                 /* data layout v3, contiguous */
            CHECK(F[b] == 3 && F[b + 1] == 1, "layout version %u class %u (expected 3, contiguous)", F[b], F[b + 1]);
            daddr = rd(b + 2, 8); dsize = rd(b + 10, 8);
            have_layout = 1;
        } else if (type == 0x0011) {                 /* symbol table: this object is a group */
            stab_btree = rd(b, 8); stab_heap = rd(b + 8, 8);
        } else {
            CHECK(type == 0x0005 || type == 0x0000, "unexpected message type 0x%04x in %s", type, path);
        }
        p = b + size;
    }
    CHECK(p == end, "messages of %s do not fill the declared header size", path);
    if (stab_btree) {
        printf("%s/ group\n", path);
        walk_group(stab_btree, stab_heap, path, leaf_k);
        return;
    }
    CHECK(have_space && have_type && have_layout, "dataset %s lacks a dataspace / datatype / layout message", path);
    uint64_t n = 1;
    for (int k = 0; k < ndim; k++) n *= dims[k];
    CHECK(dsize == n * esize, "%s: layout size %" PRIu64 " != %" PRIu64 " elements x %u bytes", path, dsize, n, esize);
    if (n) CHECK(daddr % 8 == 0 && daddr + dsize <= FSZ, "%s: raw data [%" PRIu64 ", +%" PRIu64 ") outside the file", path, daddr, dsize);
    printf("%s %s [", path, tname);
    for (int k = 0; k < ndim; k++) printf("%s%" PRIu64, k ? "," : "", dims[k]);
    printf("]");
    const uint64_t show = n < 6 ? n : 6;
    for (uint64_t i = 0; i < show; i++) {
        const uint64_t q = daddr + i * esize;
        if (!strcmp(tname, "float64")) { double v; memcpy(&v, F + q, 8); printf(" %.17g", v); }
        else if (!strcmp(tname, "float32")) { float v; memcpy(&v, F + q, 4); printf(" %.9g", v); }
        else if (!strcmp(tname, "complex128")) { double v[2]; memcpy(v, F + q, 16); printf(" (%.17g,%.17g)", v[0], v[1]); }
        else if (!strncmp(tname, "int", 3)) { int64_t v = (int64_t)rd(q, esize); if (esize < 8) { const int sh = 64 - 8 * esize; v = (int64_t)((uint64_t)v << sh) >> sh; } printf(" %" PRId64, v); }
        else printf(" %" PRIu64, rd(q, esize));
    }
    /* order-independent checksum of the raw bytes, so that the test pins every element without printing them all */
    uint64_t sum = 0;
    for (uint64_t i = 0; i < dsize; i++) sum = sum * 1099511628211ull + F[daddr + i];
    printf(" fnv=%016" PRIx64 "\n", sum);
}

static void walk_group(uint64_t btree, uint64_t heap, const char* prefix, int leaf_k) {
    CHECK(memcmp(F + heap, "HEAP", 4) == 0 && F[heap + 4] == 0, "local heap signature / version at %" PRIu64, heap);
    const uint64_t hsize = rd(heap + 8, 8), hfree = rd(heap + 16, 8), hdata = rd(heap + 24, 8);
    CHECK(hdata + hsize <= FSZ, "local heap data segment outside the file");
    CHECK(hfree == 1 || hfree + 16 <= hsize, "local heap free-list head %" PRIu64 " is not inside the data segment", hfree);
    CHECK(memcmp(F + btree, "TREE", 4) == 0 && F[btree + 4] == 0, "B-tree signature / node type at %" PRIu64, btree);
    const unsigned level = F[btree + 5], used = (unsigned)rd(btree + 6, 2);
    CHECK(level == 0, "B-tree level %u (the writer uses one leaf level)", level);
    CHECK(rd(btree + 8, 8) == 0xFFFFFFFFFFFFFFFFull && rd(btree + 16, 8) == 0xFFFFFFFFFFFFFFFFull, "B-tree sibling addresses must be undefined");
    uint64_t p = btree + 24;
    char last[512] = "";
    for (unsigned c = 0; c < used; c++) {
        const uint64_t child = rd(p + 8, 8);             /* key c (heap offset), child c, key c+1 */
        CHECK(memcmp(F + child, "SNOD", 4) == 0 && F[child + 4] == 1, "symbol-table node signature / version at %" PRIu64, child);
        const unsigned nsym = (unsigned)rd(child + 6, 2);
        CHECK(nsym <= 2u * (unsigned)leaf_k, "symbol-table node holds %u entries, more than 2 K = %d", nsym, 2 * leaf_k);
        const uint64_t key_hi = rd(p + 16, 8);
        for (unsigned s = 0; s < nsym; s++) {
            const uint64_t e = child + 8 + 40ull * s;
            const uint64_t name_off = rd(e, 8), ohdr = rd(e + 8, 8);
            CHECK(name_off < hsize, "link name offset outside the heap");
            const char* name = (const char*)F + hdata + name_off;
            CHECK(strcmp(last, name) < 0, "symbol-table entries must be sorted by name ('%s' after '%s')", name, last);
            snprintf(last, sizeof last, "%s", name);
            if (s == nsym - 1) CHECK(key_hi == name_off, "right B-tree key must be the heap offset of the node's last name");
            char path[1024];
            snprintf(path, sizeof path, "%s/%s", prefix, name);
            const unsigned cache = (unsigned)rd(e + 16, 4);
            if (cache == 1) CHECK(rd(e + 24, 8) != 0 && rd(e + 32, 8) != 0, "cached B-tree / heap addresses of group %s", path);
            object(ohdr, path, leaf_k);
        }
        p += 16;
    }
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.h5\n", argv[0]); return 2; }
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) { perror(argv[1]); return 2; }
    fseek(fp, 0, SEEK_END); FSZ = (uint64_t)ftell(fp); fseek(fp, 0, SEEK_SET);
    F = (unsigned char*)malloc(FSZ + 16);
    if (fread(F, 1, FSZ, fp) != FSZ) return 2;
    memset(F + FSZ, 0, 16);
    fclose(fp);
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    CHECK(FSZ >= 96 && memcmp(F, sig, 8) == 0, "format signature");
    CHECK(F[8] == 0 && F[9] == 0 && F[10] == 0 && F[12] == 0, "superblock / free-space / root-group / shared-header versions must be 0");
    CHECK(F[13] == 8 && F[14] == 8, "size of offsets %u / lengths %u (expected 8 / 8)", F[13], F[14]);
    const int leaf_k = (int)rd(16, 2), internal_k = (int)rd(18, 2);
    CHECK(leaf_k >= 1 && internal_k >= 1, "group leaf K %d / internal K %d", leaf_k, internal_k);
    CHECK(rd(24, 8) == 0, "base address must be 0");
    CHECK(rd(32, 8) == 0xFFFFFFFFFFFFFFFFull, "free-space info address must be undefined");
    CHECK(rd(40, 8) == FSZ, "end-of-file address %" PRIu64 " != file size %" PRIu64, rd(40, 8), FSZ);
    CHECK(rd(48, 8) == 0xFFFFFFFFFFFFFFFFull, "driver info address must be undefined");
    /* root group symbol-table entry at 56: link name offset, object header address, cache type 1 + scratch (B-tree, heap) */
    const uint64_t root_hdr = rd(64, 8);
    CHECK(rd(56, 8) == 0 && rd(72, 4) == 1, "root symbol-table entry (name offset 0, cache type 1)");
    printf("superblock v0 leafK=%d internalK=%d eof=%" PRIu64 "\n", leaf_k, internal_k, FSZ);
    object(root_hdr, "", leaf_k);
    /* the cached addresses in the superblock must be the ones the root object header's symbol-table message holds */
    {
        uint64_t p = root_hdr + 16;
        const unsigned nmsg = (unsigned)rd(root_hdr + 2, 2);
        for (unsigned m = 0; m < nmsg; m++) {
            if (rd(p, 2) == 0x0011) CHECK(rd(p + 8, 8) == rd(80, 8) && rd(p + 16, 8) == rd(88, 8), "superblock scratch-pad addresses differ from the root symbol-table message");
            p += 8 + rd(p + 2, 2);
        }
    }
    if (failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
    puts("h5_check ok");
    return 0;
}
