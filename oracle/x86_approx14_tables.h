/* Coefficients of VRCP14PS / VRSQRT14PS (oracle copy; TEST INFRASTRUCTURE).  Generated from the INSTRUCTIONS themselves by
 * tools/x86_approx14/{dump_tables.c, fit_tables.py} on an AVX-512 host: result bits = ((c0 - c1 * t) >> 9) << 7 per segment
 * (64 segments; rsqrt14: [0,32) even exponent, [32,64) odd), exact for powers of two / four.  tests/test_x86_approx14.py
 * compares the functions built on these tables with the instructions on all 2^32 inputs. */
#ifndef ORC_X86_APPROX14_TABLES_H
#define ORC_X86_APPROX14_TABLES_H
#include <stdint.h>
static const uint32_t orc_rcp14_coef[64][2] = {
    {0xFDFFF900u, 1009u}, {0xFDF03600u, 977u}, {0xFDE0F200u, 949u}, {0xFDD22000u, 921u},
    {0xFDC3BB00u, 893u}, {0xFDB5C700u, 869u}, {0xFDA83300u, 843u}, {0xFD9B0600u, 821u},
    {0xFD8E3200u, 797u}, {0xFD81BC00u, 777u}, {0xFD759800u, 755u}, {0xFD69CA00u, 735u},
    {0xFD5E4C00u, 717u}, {0xFD531B00u, 699u}, {0xFD483100u, 681u}, {0xFD3D8C00u, 663u},
    {0xFD332F00u, 647u}, {0xFD291100u, 631u}, {0xFD1F3600u, 617u}, {0xFD159300u, 601u},
    {0xFD0C2D00u, 587u}, {0xFD02FF00u, 573u}, {0xFCFA0A00u, 561u}, {0xFCF14500u, 547u},
    {0xFCE8B600u, 535u}, {0xFCE05800u, 523u}, {0xFCD82D00u, 513u}, {0xFCD02A00u, 501u},
    {0xFCC85700u, 491u}, {0xFCC0AD00u, 479u}, {0xFCB92E00u, 469u}, {0xFCB1D700u, 459u},
    {0xFCAAAA00u, 451u}, {0xFCA39F00u, 441u}, {0xFC9CBC00u, 433u}, {0xFC95F800u, 423u},
    {0xFC8F5A00u, 415u}, {0xFC88DD00u, 407u}, {0xFC828000u, 399u}, {0xFC7C4300u, 391u},
    {0xFC762800u, 385u}, {0xFC702500u, 377u}, {0xFC6A4100u, 369u}, {0xFC647B00u, 363u},
    {0xFC5ED100u, 357u}, {0xFC593D00u, 349u}, {0xFC53C600u, 343u}, {0xFC4E6800u, 337u},
    {0xFC492300u, 331u}, {0xFC43F500u, 325u}, {0xFC3EDE00u, 319u}, {0xFC39E200u, 315u},
    {0xFC34F600u, 309u}, {0xFC302100u, 303u}, {0xFC2B6400u, 299u}, {0xFC26B700u, 293u},
    {0xFC222200u, 289u}, {0xFC1D9F00u, 285u}, {0xFC192D00u, 279u}, {0xFC14D300u, 275u},
    {0xFC108900u, 271u}, {0xFC0C4F00u, 267u}, {0xFC082500u, 263u}, {0xFC040B00u, 259u},
};
static const uint32_t orc_rsqrt14_coef[64][2] = {
    {0xFDFFF480u, 1001u}, {0xFDF05080u, 955u}, {0xFDE16280u, 915u}, {0xFDD31900u, 877u},
    {0xFDC56700u, 841u}, {0xFDB84380u, 807u}, {0xFDABA680u, 775u}, {0xFD9F8880u, 747u},
    {0xFD93DD00u, 719u}, {0xFD88A080u, 693u}, {0xFD7DCB80u, 669u}, {0xFD735A00u, 647u},
    {0xFD694100u, 625u}, {0xFD5F7D00u, 603u}, {0xFD560F80u, 585u}, {0xFD4CED80u, 567u},
    {0xFD441380u, 549u}, {0xFD3B8180u, 533u}, {0xFD332F80u, 517u}, {0xFD2B1C00u, 501u},
    {0xFD234680u, 487u}, {0xFD1BA980u, 473u}, {0xFD144400u, 461u}, {0xFD0D1180u, 449u},
    {0xFD060F80u, 437u}, {0xFCFF3D80u, 425u}, {0xFCF89B00u, 415u}, {0xFCF21F00u, 403u},
    {0xFCEBCF80u, 393u}, {0xFCE5AB00u, 385u}, {0xFCDFA780u, 375u}, {0xFCD9CD00u, 367u},
    {0xFCD40A80u, 707u}, {0xFCC8FC80u, 675u}, {0xFCBE6E00u, 647u}, {0xFCB45200u, 619u},
    {0xFCAAA600u, 595u}, {0xFCA15B80u, 571u}, {0xFC987080u, 549u}, {0xFC8FDC80u, 527u},
    {0xFC879E80u, 509u}, {0xFC7FAD80u, 491u}, {0xFC780280u, 473u}, {0xFC709E80u, 457u},
    {0xFC697A80u, 441u}, {0xFC629500u, 427u}, {0xFC5BE880u, 413u}, {0xFC557580u, 401u},
    {0xFC4F3380u, 389u}, {0xFC492180u, 377u}, {0xFC433F80u, 365u}, {0xFC3D8C80u, 355u},
    {0xFC380180u, 345u}, {0xFC329F00u, 335u}, {0xFC2D6200u, 325u}, {0xFC284C00u, 317u},
    {0xFC235900u, 309u}, {0xFC1E8680u, 301u}, {0xFC19D380u, 293u}, {0xFC153F00u, 285u},
    {0xFC10CA80u, 279u}, {0xFC0C6E80u, 271u}, {0xFC083000u, 265u}, {0xFC040B00u, 259u},
};
#endif
